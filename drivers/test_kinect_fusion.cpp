// test_kinect_fusion — the YAML-driven frame-loop driver, mirroring Experiments/test_xkinect_fusion/main.cpp:16-84:
// config -> dataset -> loop { upload depth, time ProcessFrame, log slam / gt poses, optional point cloud } ->
// "mean frame time".  Host C++ over the C-ABI of libxslam_b200.so (include/xslam_b200.h); no CPU fallback.
//
// Differences from the reference driver, all stated:
//  * dataset_format: "ICL" (ICL_Dataset, as the reference driver), "7-Scenes" (seven_scenes_Dataset with the `seq_info`
//    file readInfo parses) - both through the library's OpenCV-free readers (xs_dataset_*, csrc/dataset.cpp) - or
//    "synthetic" (analytic-SDF room + closed-form trajectory, xs_synth_depth / xs_synth_pose; no datasets offline);
//  * csfd_mode (none | gradient | hessian) seeds k perturbation directions on world2camera — the batched form of the
//    commented seeding line KinectFusionReconstruction.cpp:22 — and log_pose_derivatives writes, next to every
//    frame-%06d.pose.txt, frame-%06d.dpose.txt with one row of 16 values per derivative component (d world2camera
//    / d theta, i.e. the stored h-scaled component divided by h, or by h^2 for eps1eps2);
//  * when frame alignment fails the driver stops (the reference spins forever);
//  * multi-GPU: `test_kinect_fusion cfg.yaml out/ --rank R --world W --nccl-id FILE` runs one process per GPU (device R); the
//    perturbation directions are sharded over the ranks - csfd_mode hessian: every rank carries the 6 first-order components
//    and its share of the 21 pairs - and the library all-gathers the pose records over NCCL after every frame
//    (xs_kinfu_set_comm).  Rank 0 draws the NCCL id into FILE, the other ranks wait for it; rank 0 writes the outputs, the
//    derivative log then holds the gathered components of all ranks.
#include "../include/xslam_b200.h"
#include "flat_yaml.h"

#include <algorithm>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <iomanip>
#include <vector>

int xs_driver_device_alloc(float **p, size_t floats);
int xs_driver_device_download(float *dst, const float *src, size_t floats);
void xs_driver_device_free(float *p);

static const float H_ = 1e-7f;  // Internal.h:33

// savePose, main.cpp:8-14
static void savePose(const std::string &output_dir, int frame_id, const float *pose16) {
    std::stringstream ss;
    ss << "frame-" << std::setw(6) << std::setfill('0') << frame_id << ".pose.txt";
    xs_save_pose_txt((output_dir + ss.str()).c_str(), pose16);
}

static void mul4(const float *a, const float *b, float *c) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            c[i * 4 + j] = s;
        }
}
// inverse of [A t; 0 0 0 1] by cofactors of A (Eigen's Matrix4f::inverse on dataset.getPose(0), main.cpp:72; dataset poses
// need not be exactly orthonormal)
static void affine_inverse(const float *m, float *o) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double inv[9] = {(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det,
                           (f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det,
                           (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};
    std::memset(o, 0, 16 * sizeof(float));
    for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < 3; ++cc) o[r * 4 + cc] = (float) inv[r * 3 + cc];
        o[r * 4 + 3] = (float) -(inv[r * 3] * m[3] + inv[r * 3 + 1] * m[7] + inv[r * 3 + 2] * m[11]);
    }
    o[15] = 1.f;
}
static void print4(const char *name, const float *m) {
    std::cout << name << ":\n";
    for (int i = 0; i < 4; ++i) std::cout << " " << m[i * 4] << " " << m[i * 4 + 1] << " " << m[i * 4 + 2] << " " << m[i * 4 + 3] << "\n";
}

// se3Exp generators (KinectFusionReconstruction.h:176-219, xi = [v; omega]) as row-major 4x4
static void generator(int i, float *G) {
    std::memset(G, 0, 16 * sizeof(float));
    if (i < 3) G[i * 4 + 3] = 1.f;
    if (i == 3) G[1 * 4 + 2] = -1.f, G[2 * 4 + 1] = 1.f;
    if (i == 4) G[0 * 4 + 2] = 1.f, G[2 * 4 + 0] = -1.f;
    if (i == 5) G[0 * 4 + 1] = -1.f, G[1 * 4 + 0] = 1.f;
}

int main(int argc, char *argv[]) {
    std::cout << "Demo of XKinectFusion" << std::endl;
    if (argc < 2) {
        std::cout << "please enter the config file name\n";
        return -1;
    }
    FlatYaml config(argv[1]);
    const std::string dataset_format = config.str("dataset_format");
    const int start_frame = config.i("start_frame"), end_frame = config.i("end_frame");
    std::string output_path = config.str("output_dir");
    if (argc > 2 && argv[2][0] != '-') output_path = argv[2];
    int rank = 0, world = 1;
    std::string id_file;
    for (int a = 2; a + 1 < argc; ++a) {
        if (!std::strcmp(argv[a], "--rank")) rank = std::atoi(argv[a + 1]);
        if (!std::strcmp(argv[a], "--world")) world = std::atoi(argv[a + 1]);
        if (!std::strcmp(argv[a], "--nccl-id")) id_file = argv[a + 1];
    }
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && id_file.empty())) {
        std::cerr << "usage: test_kinect_fusion cfg.yaml [out/] [--rank R --world W --nccl-id FILE]\n";
        return -1;
    }
    xs_comm *comm = nullptr;
    if (world > 1) {  // one process per GPU: device = rank; the NCCL id travels through a file
        if (xs_set_device(rank) != XS_OK) {
            std::cerr << "rank " << rank << ": " << xs_last_error() << "\n";
            return -1;
        }
        unsigned char id[128];
        if (rank == 0) {
            if (xs_comm_unique_id(id) != XS_OK) {
                std::cerr << "xs_comm_unique_id: " << xs_last_error() << "\n";
                return -1;
            }
            std::ofstream f(id_file + ".tmp", std::ios::binary);
            f.write((const char *) id, 128);
            f.close();
            std::filesystem::rename(id_file + ".tmp", id_file);
        } else {
            for (int tries = 0; !std::filesystem::exists(id_file) && tries < 600; ++tries) std::this_thread::sleep_for(std::chrono::milliseconds(100));
            std::ifstream f(id_file, std::ios::binary);
            if (!f.read((char *) id, 128)) {
                std::cerr << "rank " << rank << ": cannot read the NCCL id from " << id_file << "\n";
                return -1;
            }
        }
        comm = xs_comm_create(rank, world, id);
        if (!comm) {
            std::cerr << "xs_comm_create: " << xs_last_error() << "\n";
            return -1;
        }
    }
    if (!output_path.empty() && output_path.back() != '/') output_path += '/';
    // dataset = ICL_Dataset(dataset_dir, start_frame, end_frame, is_flip), main.cpp:34
    xs_dataset *dataset = nullptr;
    if (dataset_format == "ICL") {
        dataset = xs_dataset_open_icl(config.str("dataset_dir").c_str(), start_frame, end_frame, config.b("is_flip", false));
    } else if (dataset_format == "7-Scenes" || dataset_format == "seven_scenes") {
        int s[64], e[64];
        char names[64 * 16];
        const std::string dir = config.str("dataset_dir");
        const int n = xs_seven_scenes_read_info(config.str("seq_info", dir + "info.txt").c_str(), s, e, names, 64);
        const char *name_ptr[64];
        for (int i = 0; i < n; ++i) name_ptr[i] = names + 16 * i;
        if (n > 0) dataset = xs_dataset_open_seven_scenes(dir.c_str(), s, e, name_ptr, n, config.b("is_flip", false));
    } else if (dataset_format != "synthetic") {
        std::cerr << "dataset_format must be ICL, 7-Scenes or synthetic\n";
        return -1;
    }
    if (dataset_format != "synthetic" && !dataset) {
        std::cerr << "cannot open the dataset: " << xs_last_error() << "\n";
        return -1;
    }
    // with a dataset, frame ids index it from 0 and the loop runs to end_frame like main.cpp:45 (clamped to its size)
    const int first_frame = dataset ? 0 : start_frame;
    const int last_frame = dataset ? std::min(end_frame, xs_dataset_size(dataset)) : end_frame - start_frame;
    std::cout << "frame num: " << (dataset ? xs_dataset_size(dataset) : end_frame - start_frame) << std::endl;
    std::cout << "initialize kinect fusion......" << std::endl;
    // KinectFusionReconstruction::SetYamlParameters, KinectFusionReconstruction.cpp:12-72
    xs_config cfg;
    cfg.res[0] = config.i("tsdf_size_x"), cfg.res[1] = config.i("tsdf_size_y"), cfg.res[2] = config.i("tsdf_size_z");
    cfg.voxel_size = config.f("tsdf_voxel_size");
    cfg.max_weight = config.i("max_integration_weight");
    cfg.thres_range = config.f("thres_range");
    cfg.init_xyz[0] = config.f("init_x"), cfg.init_xyz[1] = config.f("init_y"), cfg.init_xyz[2] = config.f("init_z");
    cfg.r_deg[0] = config.f("r_x"), cfg.r_deg[1] = config.f("r_y"), cfg.r_deg[2] = config.f("r_z");
    cfg.width = config.i("depth_width"), cfg.height = config.i("depth_height");
    cfg.fx = config.f("fx"), cfg.fy = config.f("fy"), cfg.cx = config.f("cx"), cfg.cy = config.f("cy");
    cfg.num_levels = config.i("num_levels");
    if (cfg.num_levels > 3) std::cout << "sorry, the max supported multi-level = 3" << std::endl;
    cfg.dist_thres = config.f("distThres");
    cfg.angle_thres_deg = config.f("angleThres");
    cfg.bi_threshold = config.f("biInterpolate_threshold");
    cfg.trunc_k = config.f("trunc_logistic_k");
    cfg.frame_step = config.i("frame_step", 1);  // KinectFusionReconstruction.cpp:72: frame_id += frame_step per processed frame
    if (cfg.frame_step < 1) {
        std::cerr << "frame_step must be >= 1\n";
        return -1;
    }
    const xs_intr intr = {cfg.fx, cfg.fy, cfg.cx, cfg.cy};

    // perturbation directions, sharded over the ranks (round robin)
    const std::string mode = config.str("csfd_mode", "none");
    int comps = 1, dirs = 0, ncomp = 0, ncomp_max = 0, nparams = 0;
    std::vector<float> seeds;
    std::vector<int> my_pairs;  // hessian mode: (i, j) of this rank's share
    xs_kinfu *kinfu = nullptr;
    if (mode == "gradient") {
        comps = 1;
        for (int i = rank; i < 6; i += world) {
            float G[16];
            generator(i, G);
            for (int e = 0; e < 16; ++e) seeds.push_back(H_ * G[e]);
            ++dirs;
        }
        ncomp = dirs;
        ncomp_max = (6 + world - 1) / world;
        kinfu = xs_kinfu_create(&cfg, 1, dirs, seeds.empty() ? nullptr : seeds.data(), XS_SOLVE_EIGEN_LLT);
    } else if (mode == "hessian") {
        // Hessian batch over the 6 pose parameters: first-order seeds h G_i, second-order seeds h^2 (G_i G_j + G_j G_i) / 2
        comps = 2, nparams = dirs = 6;
        for (int i = 0; i < 6; ++i) {
            float G[16];
            generator(i, G);
            for (int e = 0; e < 16; ++e) seeds.push_back(H_ * G[e]);
        }
        int k = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j, ++k) {
                if (k % world != rank) continue;
                float Gi[16], Gj[16], GiGj[16], GjGi[16];
                generator(i, Gi), generator(j, Gj);
                mul4(Gi, Gj, GiGj), mul4(Gj, Gi, GjGi);
                for (int e = 0; e < 16; ++e) seeds.push_back(H_ * H_ * 0.5f * (GiGj[e] + GjGi[e]));
                my_pairs.push_back(i);
                my_pairs.push_back(j);
            }
        ncomp = 6 + (int) my_pairs.size() / 2;
        ncomp_max = 6 + (21 + world - 1) / world;
        kinfu = xs_kinfu_create_hessian(&cfg, 6, (int) my_pairs.size() / 2, my_pairs.data(), seeds.data(), XS_SOLVE_ANALYTIC);
    } else if (mode == "none") {
        kinfu = xs_kinfu_create(&cfg, 1, 0, nullptr, XS_SOLVE_EIGEN_LLT);
    } else {
        std::cerr << "csfd_mode must be none, gradient or hessian\n";
        return -1;
    }
    if (!kinfu) {
        std::cerr << "initialisation failed: " << xs_last_error() << "\n";
        return -1;
    }
    const int record_floats = (1 + ncomp_max) * 16;
    if (comm && xs_kinfu_set_comm(kinfu, comm, record_floats) != XS_OK) {
        std::cerr << "xs_kinfu_set_comm: " << xs_last_error() << "\n";
        return -1;
    }
    const bool writer = rank == 0;  // rank 0 writes the outputs
    const bool log_slam = config.b("log_slam_pose") && writer, log_gt = config.b("log_gt_pose") && writer, draw_pcd = config.b("draw_pcd") && writer;
    const bool log_deriv = config.b("log_pose_derivatives", false) && mode != "none" && writer;
    std::vector<float> gathered((size_t) world * record_floats);
    double total_time = 0;
    std::cout << "start slam!" << std::endl;
    std::vector<uint16_t> depth((size_t) cfg.width * cfg.height);
    std::vector<float> w2c((size_t) (1 + ncomp) * 16);
    float gt0_inv[16];
    {
        float gt0[16];
        if (dataset)
            xs_dataset_get_pose(dataset, 0, gt0);
        else
            xs_synth_pose(start_frame, gt0);
        affine_inverse(gt0, gt0_inv);
    }
    std::vector<float> pts, nrm;
    while (xs_kinfu_frame_id(kinfu) < last_frame) {
        const int frame_id = xs_kinfu_frame_id(kinfu);
        std::cout << "current frame is " << frame_id << "\n";
        float gt_pose[16];
        if (dataset) {  // dataset.getDepthData(frame_id, depth_map), main.cpp:50
            if (xs_dataset_get_depth(dataset, first_frame + frame_id, depth.data(), cfg.height, cfg.width) != XS_OK) {
                std::cerr << "getDepthData failed: " << xs_last_error() << "\n";
                return -1;
            }
            xs_dataset_get_pose(dataset, first_frame + frame_id, gt_pose);
        } else {
            xs_synth_pose(start_frame + frame_id, gt_pose);
            xs_synth_depth(gt_pose, intr, cfg.height, cfg.width, depth.data());
        }
        // c. process kinect fusion (timed like main.cpp:57-60; the upload of the frame is inside ProcessFrame here)
        const auto t0 = std::chrono::steady_clock::now();
        const int ok = xs_kinfu_process_frame(kinfu, depth.data(), 0);
        const double frame_time = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (!ok) {
            std::cerr << "Frame align failed! (" << xs_last_error() << ")\n";
            break;
        }
        total_time += frame_time;
        float pose_c2w[16];
        xs_kinfu_get_pose_c2w(kinfu, pose_c2w);
        if (log_slam) {
            std::filesystem::create_directories(output_path + "slam/");
            print4("slam c2w", pose_c2w);
            savePose(output_path + "slam/", frame_id, pose_c2w);
        }
        if (log_gt) {
            std::filesystem::create_directories(output_path + "gt/");
            float gt_c2w[16];
            mul4(gt0_inv, gt_pose, gt_c2w);  // dataset.getPose(0).inverse() * dataset.getPose(frame_id), main.cpp:72
            print4("gt c2w", gt_c2w);
            savePose(output_path + "gt/", frame_id, gt_c2w);
        }
        if (log_deriv) {
            // one row per derivative component of world2camera, unscaled (d / d theta_i, then d2 / d theta_i d theta_j in the
            // order of the pair list): this rank's own record, or with several ranks the records the library gathered
            std::filesystem::create_directories(output_path + "slam/");
            std::stringstream ss;
            ss << output_path << "slam/frame-" << std::setw(6) << std::setfill('0') << frame_id << ".dpose.txt";
            std::ofstream out(ss.str());
            auto row = [&](const float *m, double scale) {
                for (int e = 0; e < 16; ++e) out << std::setprecision(7) << std::scientific << m[e] * scale << " ";
                out << "\n";
            };
            if (!comm) {
                xs_kinfu_get_world2camera(kinfu, w2c.data());
                for (int q = 0; q < ncomp; ++q) row(&w2c[(size_t) (1 + q) * 16], comps == 2 && q >= nparams ? 1.0 / ((double) H_ * H_) : 1.0 / H_);
            } else {
                xs_kinfu_get_gathered_records(kinfu, gathered.data());
                if (comps == 2) {
                    for (int i = 0; i < 6; ++i) row(&gathered[(size_t) (1 + i) * 16], 1.0 / H_);
                    for (int k = 0; k < 21; ++k) row(&gathered[(size_t) (k % world) * record_floats + (size_t) (1 + 6 + k / world) * 16], 1.0 / ((double) H_ * H_));
                } else {
                    for (int i = 0; i < 6; ++i) row(&gathered[(size_t) (i % world) * record_floats + (size_t) (1 + i / world) * 16], 1.0 / H_);
                }
            }
        }
        if (draw_pcd && frame_id + cfg.frame_step >= last_frame) {
            // ExportPointCloud(1000000) + exportPly on the last frame, main.cpp:76-81 / KinectFusionReconstruction.cpp:334-372
            const long max_buffer = 1000000;
            float *d_pts = nullptr, *d_nrm = nullptr;
            if (xs_driver_device_alloc(&d_pts, 3 * max_buffer) || xs_driver_device_alloc(&d_nrm, 3 * max_buffer)) {
                std::cerr << "point-cloud buffers: out of device memory\n";
                return -1;
            }
            const long n = xs_extract_points(xs_kinfu_volume(kinfu), d_pts, d_nrm, max_buffer, nullptr);
            if (n < 0) {
                std::cerr << "ExportPointCloud failed: " << xs_last_error() << "\n";
                return -1;
            }
            pts.resize((size_t) 3 * n), nrm.resize((size_t) 3 * n);
            xs_driver_device_download(pts.data(), d_pts, (size_t) 3 * n);
            xs_driver_device_download(nrm.data(), d_nrm, (size_t) 3 * n);
            xs_driver_device_free(d_pts), xs_driver_device_free(d_nrm);
            xs_export_ply((output_path + "pcd.ply").c_str(), pts.data(), nrm.data(), n);
            std::cout << "point cloud: " << n << " points -> " << output_path << "pcd.ply\n";
        }
    }
    const int frames = (xs_kinfu_frame_id(kinfu) + cfg.frame_step - 1) / cfg.frame_step;  // frames processed
    printf("mean frame time = %.3f ms\n", total_time / (frames > 0 ? frames : 1));
    xs_kinfu_destroy(kinfu);
    xs_comm_destroy(comm);
    xs_dataset_close(dataset);
    return 0;
}
