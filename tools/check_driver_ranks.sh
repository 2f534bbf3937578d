#!/bin/bash
# C++ driver on 2 ranks (one process per GPU, NCCL id through a file) against the 1-rank run: hessian mode, 4 frames;
# the gathered derivative log of rank 0 must equal the 1-rank log.  usage: tools/check_driver_ranks.sh <outdir>
OUT=${1:-gpurun_out/driver_ranks}
rm -rf $OUT; mkdir -p $OUT
sed -e 's/end_frame: 30/end_frame: 4/' -e 's/csfd_mode: gradient/csfd_mode: hessian/' -e 's/draw_pcd: true/draw_pcd: false/' configs/synth_traj2.yaml > $OUT/cfg.yaml
grep -q "csfd_mode: hessian" $OUT/cfg.yaml || echo "csfd_mode: hessian" >> $OUT/cfg.yaml
grep -q "log_pose_derivatives" $OUT/cfg.yaml || echo "log_pose_derivatives: true" >> $OUT/cfg.yaml
BIN=x-slam_b200/bin/test_kinect_fusion
$BIN $OUT/cfg.yaml $OUT/n1/ > $OUT/n1.log 2>&1 || { echo "1-rank driver failed"; tail -5 $OUT/n1.log; exit 1; }
rm -f $OUT/nccl.id
$BIN $OUT/cfg.yaml $OUT/n2/ --rank 1 --world 2 --nccl-id $OUT/nccl.id > $OUT/n2_r1.log 2>&1 &
P1=$!
$BIN $OUT/cfg.yaml $OUT/n2/ --rank 0 --world 2 --nccl-id $OUT/nccl.id > $OUT/n2_r0.log 2>&1 || { echo "rank 0 failed"; tail -5 $OUT/n2_r0.log; kill $P1; exit 1; }
wait $P1 || { echo "rank 1 failed"; tail -5 $OUT/n2_r1.log; exit 1; }
python - "$OUT" <<'PY'
import sys, glob, numpy as np
out = sys.argv[1]
worst = 0.0
files = sorted(glob.glob(out + "/n1/slam/frame-*.dpose.txt"))
assert len(files) == 4, files
for f in files:
    a = np.loadtxt(f)
    b = np.loadtxt(f.replace("/n1/", "/n2/"))
    assert a.shape == b.shape == (27, 16), (a.shape, b.shape)
    p1 = open(f.replace(".dpose", ".pose")).read()
    p2 = open(f.replace("/n1/", "/n2/").replace(".dpose", ".pose")).read()
    assert p1 == p2, "real poses differ between 1 and 2 ranks"
    for r in range(27):
        sc = max(np.abs(a[r]).max(), 1e-12)
        worst = max(worst, float(np.abs(a[r] - b[r]).max() / sc))
print("driver 2-rank vs 1-rank: real poses identical, worst derivative row rel diff %.3g" % worst)
assert worst <= 1e-5
PY
