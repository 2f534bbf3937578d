// flat_yaml.h — reader for the flat `key: value` YAML files the reference drivers use
// (Experiments/test_xkinect_fusion/configs/ICL_traj2.yaml).  yaml-cpp is not available offline; the reference config
// has no nesting, sequences or anchors, so a line reader covers it: comments (#), quoted strings, bool / int / float.
#pragma once
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>

class FlatYaml {
  public:
    explicit FlatYaml(const std::string &path) {
        std::ifstream in(path);
        if (!in) throw std::runtime_error("cannot open config file " + path);
        std::string line;
        while (std::getline(in, line)) {
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            const size_t colon = line.find(':');
            if (colon == std::string::npos) continue;
            std::string k = trim(line.substr(0, colon)), v = trim(line.substr(colon + 1));
            if (v.size() >= 2 && (v.front() == '"' || v.front() == '\'') && v.back() == v.front()) v = v.substr(1, v.size() - 2);
            if (!k.empty()) kv_[k] = v;
        }
    }
    bool has(const std::string &k) const { return kv_.count(k) != 0; }
    std::string str(const std::string &k) const {
        auto it = kv_.find(k);
        if (it == kv_.end()) throw std::runtime_error("config key missing: " + k);  // yaml-cpp throws on a missing key too
        return it->second;
    }
    std::string str(const std::string &k, const std::string &dflt) const { return has(k) ? kv_.at(k) : dflt; }
    int i(const std::string &k) const { return std::stoi(str(k)); }
    int i(const std::string &k, int dflt) const { return has(k) ? i(k) : dflt; }
    float f(const std::string &k) const { return std::stof(str(k)); }
    bool b(const std::string &k) const {
        const std::string v = str(k);
        return v == "true" || v == "True" || v == "1";
    }
    bool b(const std::string &k, bool dflt) const { return has(k) ? b(k) : dflt; }

  private:
    static std::string trim(const std::string &s) {
        const size_t a = s.find_first_not_of(" \t\r\n"), z = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, z - a + 1);
    }
    std::map<std::string, std::string> kv_;
};
