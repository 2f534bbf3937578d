#!/bin/bash
# Copies a round capture (tools/gpu_round.sh TAG) from gpurun_out/ into profiles/ and writes its ncu summary.  usage: tools/collect_profiles.sh TAG
TAG=$1
cp gpurun_out/bench_$TAG.json profiles/${TAG}_bench.json
cp gpurun_out/bench_ref_$TAG.json profiles/${TAG}_bench_reference_arm.json
cp gpurun_out/launches_$TAG.csv profiles/${TAG}_launches.csv
cp gpurun_out/test_$TAG.log profiles/${TAG}_pytest_gpu.log
for f in parity_report pipeline_csfd_report pipeline_dcsfd_report bench_config_parity_512 intrinsics_pipeline_vs_oracle intrinsics_surface intrinsics_raycast_fd hessian_batch_stages hessian_batch_pipeline_all_pairs hessian_batch_pipeline_pair_subset hessian_blocked_shards second_order_vs_dual_oracle seam_run dc_array_report; do
  [ -f gpurun_out/$f.json ] && cp gpurun_out/$f.json profiles/${TAG}_$f.json
done
cp gpurun_out/fd_second_order.txt profiles/${TAG}_fd_second_order.txt
{
echo "# Round 2 - ncu captures of the default bench workload (capture $TAG)"; echo
echo "Command: \`python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda\` under \`ncu --set full --clock-control none --import-source on\`, one kernel instance of frame 2 each (tools/gpu_ncu.sh); launch list: \`ncu --metrics gpu__time_duration.sum --clock-control none\` of the same command with 4 frames per step (tools/gpu_round.sh, profiles/${TAG}_launches.csv).  Workload: 640x480, 512^3, Hessian batch of 10 parameters (6 pose DoF + fx, fy, cx, cy), 55 pairs = 65 derivative planes.  Times under ncu are cold-cache and serialised; the bench line (profiles/${TAG}_bench.json) holds the CUDA-event times."; echo
echo "## Per-kernel metrics (--set full)"; echo
for k in icp raycast integrate; do python tools/ncu_summary.py rep gpurun_out/prof_${TAG}_$k.ncu-rep; echo; done
echo "## Launch list (12 frames: shares of the step)"; echo
python tools/ncu_summary.py list gpurun_out/launches_$TAG.csv; echo
echo "\`reset_volume_kernel\` runs once at start-up (volume clear), outside the timed region."; echo
echo "## Stall reasons of icp_deriv_tile_kernel (warps per issue)"; echo; echo '```'
ncu -i gpurun_out/prof_${TAG}_icp.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; r=rows[2]
for i,n in enumerate(h):
    if 'issue_stalled' in n and 'ratio' in n and 'not_issued' not in n:
        try:
            v=float(r[i])
            if v>0.05: print('%-40s %.2f'%(n.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v))
        except: pass
"
echo '```'; echo; echo "## Source-level hot spots (share of warp instructions / of stall samples)"; echo
for k in icp:icp_deriv_tile raycast:raycast_hit integrate:integrate_kernel; do echo "### ${k#*:}"; echo '```'; python tools/ncu_source.py gpurun_out/prof_${TAG}_${k%%:*}.ncu-rep ${k#*:} 14 --by-inst | tail -15 | cut -c1-170; echo '```'; done
} > profiles/${TAG}_ncu_summary.md
cuobjdump -sass -fun '_ZN2xs21icp_deriv_tile_kernelILi11ELi5ELb1ELb1ELi1EEEvNS_9IcpParamsENS_11SolveParamsE' x-slam_b200/libxslam_b200.so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//' | cut -c1-110 > /tmp/tile_full.sass
A=$(grep -n "DEPBAR" /tmp/tile_full.sass | sed -n 2p | cut -d: -f1); B=$(grep -n "DEPBAR" /tmp/tile_full.sass | tail -1 | cut -d: -f1)
{ echo "# SASS of icp_deriv_tile_kernel<11, 5, CURR, PIPE, DEPTH = 1> (sm_100a, cuobjdump -sass), main loop only: from the per-tile cp.async wait + barrier to the loop's back edge"; sed -n "$((A-10)),$((B+5))p" /tmp/tile_full.sass; } > profiles/${TAG}_sass_icp_deriv_tile_kernel.txt
ls profiles | grep "^$TAG" | wc -l
