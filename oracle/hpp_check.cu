// oracle/hpp_check.cu — TEST INFRASTRUCTURE ONLY: compile check of include/xslam_b200.hpp against the REFERENCE's own
// types.  Instantiates every seam wrapper with DeviceArray2D / PtrStep / Intr / MatS33 / devComplex3 from
// /root/reference (XKinectFusion/include/Internal.h, DeviceArray/include/*.hpp), i.e. with exactly the argument types
// KinectFusionReconstruction.cpp passes (:143,198,264,274-275,290,293-296,327).  Built by `make -C oracle ref`
// (object file only; nothing links or runs it).
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
}
