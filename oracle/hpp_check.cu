// oracle/hpp_check.cu — TEST INFRASTRUCTURE ONLY: compile check of include/xslam_b200.hpp against the REFERENCE's own
// types.  Instantiates every seam wrapper with DeviceArray2D / PtrStep / Intr / MatS33 / devComplex3 from
// /root/reference (XKinectFusion/include/Internal.h, DeviceArray/include/*.hpp), i.e. with exactly the argument types
// KinectFusionReconstruction.cpp passes (:143,198,264,274-275,290,293-296,327).  Built by `make -C oracle ref`
// (object file only: this translation unit checks that every wrapper INSTANTIATES with the reference types, including the
// dormant operators; oracle/seam_run.cu is the linked, executed counterpart for the frame loop's call sequence).
#include "Internal.h"
#include "../include/xslam_b200.hpp"

namespace S = xslam_b200::seam;

void hpp_check_instantiate() {
    DeviceArray2D<ushort> depth_raw;
    MapArr depth0, depth1, vmap, nmap, vmap1, nmap1, vprev, nprev;
    Intr intr(481.2f, -480.f, 319.5f, 239.5f);
    S::bilateralFilter(depth_raw, depth0);          // Map.h:16
    S::pyrDown(depth0, depth1);                     // Map.h:22
    S::createVMap(intr(0), depth0, vmap);           // Map.h:29
    S::createNMap(vmap, nmap);                      // Map.h:35
    S::resizeVMap(vmap, vmap1);                     // Map.h:47
    S::resizeNMap(nmap, nmap1);                     // Map.h:54
    MatS33 R;
    devComplex3 t;
    int3 res = make_int3(256, 256, 256);
    DeviceArray2D<float> value, grad, depthScaled;
    DeviceArray2D<int> weight;
    PtrStepSz<ushort> d = depth_raw;
    S::integrateTsdfVolume(d, intr, 100, res, 0.03f, R, t, t, 0.09f, (PtrStep<float>) value, (PtrStep<int>) weight,
                           (PtrStep<float>) grad, depthScaled, 0, 0.f, 0.f);  // TsdfFusion.h:40-45
    S::raycast(intr, R, t, R, t, 0.09f, res, 0.03f, (PtrStep<float>) value, (PtrStep<float>) grad, vprev, nprev);  // RayCaster.h:21-25
    DeviceArray2D<devComplexICP> gbuf;
    DeviceArray<devComplexICP> mbuf;
    hostComplexICP A[36], b[6];
    S::estimateCombined(R, t, vmap, nmap, R, t, intr, vprev, nprev, 0.1f, 0.26f, gbuf, mbuf, A, b);  // ICP.h:24-31
    DeviceArray2D<float> jacobi_buf, hessian_buf[12];
    Eigen::Matrix4f jacobi_host, hessian_store[3][4], *hessian_host[3] = {hessian_store[0], hessian_store[1], hessian_store[2]};
    S::computeOptimizeMatrix(vmap, nmap, vprev, nprev, R, t, R, t, intr, 0.1f, 0.26f, jacobi_buf, jacobi_host, hessian_buf,
                             hessian_host);  // ICP.h:34-40
    thrustDvec<float> gt_vec, real_vec, grad_vec, hessian_vec;
    thrustDvec<int> count_vec;
    MatD33 RD;
    devDComplex3 tD;
    float4 h4 = S::ComputeLocalTsdf_hessian(d, intr, depthScaled, res, 0.03f, RD, tD, 0.09f, 0.f, 0.f, gt_vec, real_vec, grad_vec,
                                            hessian_vec, count_vec);  // TsdfFusion.h:55-60
    Mat33 Rf;
    float3 tf = make_float3(0.f, 0.f, 0.f);
    float2 l2 = S::ComputeLocalTsdf_loss(d, intr, depthScaled, res, 0.03f, Rf, tf, 0.09f, 0.f, 0.f, gt_vec, real_vec,
                                         count_vec);  // TsdfFusion.h:48-52
    DeviceArray2D<int> packed;
    S::initVolume(packed, value, weight, grad, res);  // TsdfVolume.h:16, as TsdfVolume::reset calls it (TsdfVolume.cpp:50)
    DeviceArray<float3> cloud, normals;
    const size_t n_pts = S::extractPoints(value, weight, grad, res, 0.03f, cloud);  // ExtractPointCloud.h:19-20
    S::extractNormals(value, weight, grad, res, 0.03f, cloud, normals);               // ExtractPointCloud.h:22-23
    (void) n_pts;
    (void) h4;
    (void) l2;
}
