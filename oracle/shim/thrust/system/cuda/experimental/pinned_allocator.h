// Stand-in for the header removed in CUDA 12 (cx.h:29 only names the type in an
// alias template).  TEST INFRASTRUCTURE ONLY — see oracle/shim/Eigen/Dense.
#pragma once
#include <memory>
namespace thrust { namespace cuda { namespace experimental {
template <class T> using pinned_allocator = std::allocator<T>;
}}}
