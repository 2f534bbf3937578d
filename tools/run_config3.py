#!/usr/bin/env python
"""BASELINE.json configs[2] as written: ICL-NUIM traj2-shaped synthetic 300-frame sequence, 640x480, 512^3 TSDF, per-frame 6-DoF
CSFD gradient of the pose, on N = 1 / 2 / 4 / 8 B200 (torchrun; N = 1 may be run plainly).  The 6 first-order directions are
sharded over the ranks (ranks beyond the sixth carry the real state only), the library all-gathers the pose records every
frame.  Reports differentiated frames/s (CUDA events over the sequence, max over ranks), the drift of the estimated trajectory
against the generating (ground-truth) trajectory, and - on one GPU - against the zero-seed run of the REFERENCE's own kernels
(oracle/_ref/libxslam_ref.so) over the same 300 frames.  One JSON line; also written to gpurun_out/config3_n<N>.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import xslam_b200 as xs
    from xslam_b200 import parallel
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    frames = int(os.environ.get("XS_FRAMES", "300"))
    res = int(os.environ.get("XS_RES", "512"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=res, tsdf_size_y=res, tsdf_size_z=res, tsdf_voxel_size=7.68 / res)
    seeds = xs.pose_seeds_csfd()
    mine = list(range(rank, 6, world))
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=1, seeds=seeds[mine] if mine else None)
    L = (1 + (6 + world - 1) // world) * 16
    if world > 1:
        comm = parallel.Comm.from_torch_distributed(dist, device="cuda")
        k.set_comm(comm, L)
    k.set_deferred(True)
    depth = [xs.synth_depth(f) for f in range(frames)]
    dev = [torch.from_numpy(d.astype(np.int16)).cuda() for d in depth]
    gt = np.stack([xs.synth_pose(f) for f in range(frames)]).astype(np.float64)
    stream = torch.cuda.ExternalStream(k.stream_ptr())
    poses, grads = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k.ProcessFrame(dev[0])  # frame 0 has no ICP (main.cpp times frames 1..)
    k.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    t0 = time.perf_counter()
    for f in range(1, frames):
        if k.ProcessFrame(dev[f]) != 1:
            raise SystemExit("frame %d: alignment failed" % f)
        poses.append(k.pose_c2w().astype(np.float64))
        if world > 1:
            grads.append(parallel.assemble_list_records(k.gathered_records(), 6, 1, world))
        else:
            grads.append(k.world2camera.reshape(-1, 16).copy())
    e1.record(stream)
    k.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t_dev = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt[0])
    if rank == 0:
        poses = np.stack(poses)
        terr = np.linalg.norm(poses[:, :3, 3] - gt[1:, :3, 3], axis=1)
        rerr = np.degrees(np.arccos(np.clip((np.einsum("fij,fij->f", poses[:, :3, :3], gt[1:, :3, :3]) - 1) / 2, -1, 1)))
        path = float(np.linalg.norm(np.diff(gt[:, :3, 3], axis=0), axis=1).sum())
        g = np.stack(grads)  # [frames - 1, 7, 16] world2camera + d/d(xi_i), h-scaled
        line = {"config": "configs[2]: %d-frame synthetic ICL-shaped sequence, 640x480, %d^3 TSDF, 6-DoF CSFD gradient per frame" % (frames, res),
                "n_gpus": world, "frames": frames, "metric": "differentiated_frames_per_s", "value": (frames - 1) / t_dev, "ms_per_frame": t_dev / (frames - 1) * 1e3,
                "wall_fps_with_host_reads": (frames - 1) / wall, "directions": 6, "directions_rank0": len(mine),
                "drift_vs_ground_truth": {"translation_rmse_m": float(np.sqrt((terr ** 2).mean())), "translation_final_m": float(terr[-1]),
                                          "translation_max_m": float(terr.max()), "rotation_final_deg": float(rerr[-1]), "rotation_max_deg": float(rerr.max()),
                                          "path_length_m": path, "final_drift_percent_of_path": float(100 * terr[-1] / path)},
                "gradient_norm_last_frame": [float(np.abs(g[-1, 1 + i]).max() / xs.H_) for i in range(6)],
                "gradient_finite": bool(np.isfinite(g).all())}
        if world == 1 and os.environ.get("XS_REF", "1") == "1":
            from oracle import pyref
            if os.path.exists(pyref.REF_CUDA_PATH):
                del k
                torch.cuda.empty_cache()
                ref = pyref.RefCuda().kinfu(cfg, None)
                rp = []
                t0 = time.perf_counter()
                for f in range(frames):
                    assert ref.process_frame(depth[f]) == 1
                    if f > 0:
                        rp.append(np.linalg.inv(ref.pose().real.astype(np.float64)))
                tref = time.perf_counter() - t0
                rp = np.stack(rp)
                d = np.abs(poses - rp).max(axis=(1, 2))
                line["vs_reference_kernels_zero_seed"] = {"pose_abs_max_over_sequence": float(d.max()), "pose_abs_final": float(d[-1]), "pose_abs_frame_10": float(d[9]),
                                                          "reference_frames_per_s_one_pass": frames / tref,
                                                          "reference_drift_final_m": float(np.linalg.norm(rp[-1, :3, 3] - gt[-1, :3, 3]))}
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "config3_n%d.json" % world), "w") as fh:
            json.dump(line, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
