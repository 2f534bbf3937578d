#!/usr/bin/env python
"""BASELINE.json configs[4] as written: 7-Scenes-shaped large volume, 1024^3 TSDF, relocalisation-style derivatives w.r.t. a
multi-frame pose set (>= 64 directions), on the GPUs of one box (torchrun, N ranks).

  part A  the frame loop at 1024^3 with 64 first-order (CSFD) pose-space directions sharded over the ranks (8 per rank at N = 8:
          (2 + 8) planes x 4 GiB = 40 GiB per GPU), records all-gathered by the library every frame: differentiated frames/s;
  part B  the relocalisation loss of the reference's todo list (ComputeLocalTsdf_hessian, TsdfFusion.cu:286-331) over a pose set
          of F frames, each parameterised by se3Exp (KinectFusionReconstruction.h:176-219): v2c_f(xi_f) = se3Exp(xi_f) v2c_f,
          all 21 parameter pairs of every frame = 21 F bicomplex directions (84 for F = 4), sharded over the ranks, evaluated
          with xs_tsdf_hessian_batch against the mapped volume as ground truth, loss rows all-gathered with the library's
          communicator; rank 0 assembles the per-frame gradient and 6x6 Hessian and checks them (the gradient of parameter i is
          the same in every pair it occurs in; one row against the single-direction entry point).
One JSON line; also written to gpurun_out/config5_n<N>.json.  XS_RES / XS_DIRS / XS_FRAMES shrink it for small boxes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ctypes as C
    import torch
    import xslam_b200 as xs
    from xslam_b200 import ops, parallel
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    res = int(os.environ.get("XS_RES", "1024"))
    ndirs = int(os.environ.get("XS_DIRS", "64"))
    frames = int(os.environ.get("XS_FRAMES", "12"))
    nset = int(os.environ.get("XS_POSE_SET", "4"))
    torch.cuda.set_device(local)
    dist, comm = None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = parallel.Comm.from_torch_distributed(dist, device="cuda")
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=res, tsdf_size_y=res, tsdf_size_z=res, tsdf_voxel_size=7.68 / res)
    # ---------------- part A: frame loop, ndirs first-order directions (6 axes + mixed pose-space directions)
    G = xs.se3_generators().reshape(6, 16)
    rng = np.random.default_rng(11)
    Wm = np.concatenate([np.eye(6), rng.standard_normal((max(ndirs - 6, 0), 6)) / np.sqrt(6)])[:ndirs]
    seeds = (xs.H_ * Wm @ G).astype(np.float32)
    mine = list(range(rank, ndirs, world))
    need = (2 + len(mine)) * res ** 3 * 4 / 2 ** 30
    free = torch.cuda.mem_get_info()[0] / 2 ** 30
    if need > 0.9 * free:
        raise SystemExit("rank %d: %d planes at %d^3 need %.0f GiB, %.0f GiB free: use more ranks or fewer directions" % (rank, 2 + len(mine), res, need, free))
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=1, seeds=seeds[mine])
    L = (1 + (ndirs + world - 1) // world) * 16
    if comm is not None:
        k.set_comm(comm, L)
    k.set_deferred(True)
    depth = [xs.synth_depth(f) for f in range(frames)]
    dev = [torch.from_numpy(d.astype(np.int16)).cuda() for d in depth]
    stream = torch.cuda.ExternalStream(k.stream_ptr())
    k.ProcessFrame(dev[0])
    k.ProcessFrame(dev[1])  # warm-up of the ICP path
    k.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    rec = None
    for f in range(2, frames):
        if k.ProcessFrame(dev[f]) != 1:
            raise SystemExit("frame %d: alignment failed" % f)
        rec = parallel.assemble_list_records(k.gathered_records(), ndirs, 1, world) if comm is not None else k.world2camera.reshape(-1, 16)
    e1.record(stream)
    k.sync()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt[0])
    tm, _ = k.times()
    vol_gib = xs.load().xs_volume_bytes(xs.load().xs_kinfu_volume(k.h)) / 2 ** 30
    part_a = {"frames_timed": frames - 2, "value": (frames - 2) / t_dev, "unit": "frames/s", "ms_per_frame": t_dev / (frames - 2) * 1e3, "directions": ndirs,
              "directions_rank0": len(mine), "volume_GiB_rank0": vol_gib, "stages_ms_last_frame": tm,
              "gathered_record_rows": int(rec.shape[0]), "derivatives_finite": bool(np.isfinite(rec).all()),
              "derivative_norms_first_6": [float(np.abs(rec[1 + i]).max() / xs.H_) for i in range(min(6, ndirs))]}
    # ---------------- part B: relocalisation Hessian over a pose set, against the mapped volume
    gt, _, _ = k.volume_planes(0)  # dense [z, y, x] value plane of the map = the ground-truth TSDF of the loss
    trunc = xs.load().xs_volume_trunc_dist(xs.load().xs_kinfu_volume(k.h))
    del k
    torch.cuda.empty_cache()
    intr = xs.Intr(cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])
    pairs = xs.all_pairs(6)
    all_dirs = [(f, p) for f in range(nset) for p in range(len(pairs))]  # (frame of the pose set, pair)
    my_dirs = all_dirs[rank::world]
    w2v = np.eye(4)
    w2v[:3, 3] = [cfg["init_x"], cfg["init_y"], cfg["init_z"]]
    set_frames = [2 + f * max(1, (frames - 3) // max(nset - 1, 1)) for f in range(nset)]
    rows_local = np.zeros((len(all_dirs[0::world]), 4), np.float64)  # padded to the largest share
    t0 = time.perf_counter()
    for f in range(nset):
        mine_f = [(i, p) for i, (ff, p) in enumerate(my_dirs) if ff == f]
        if not mine_f:
            continue
        c2w = xs.synth_pose(set_frames[f]).astype(np.float64)
        v2c0 = np.linalg.inv(w2v @ c2w)
        # xi_f = 0 with DCSFD seeds: direction (i, j) perturbs xi along e_i (eps1) and e_j (eps2); se3Exp carries them through
        dR, dt = [], []
        for _, p in mine_f:
            i, j = pairs[p]
            xi = np.zeros((4, 6), np.float32)
            xi[1, i] = xs.H_
            xi[2, j] = xs.H_
            T = xs.se3_exp(xi, comps=3, dirs=1).astype(np.float64)  # [4, 4, 4]: real, eps1, eps2, eps1eps2
            for c in range(1, 4):
                M = T[c] @ v2c0  # d(se3Exp(xi) v2c_f)
                dR.append(M[:3, :3].reshape(9))
                dt.append(M[:3, 3])
        pb = ops.PoseBatch(v2c0[:3, :3], v2c0[:3, 3], np.asarray(dR, np.float32), np.asarray(dt, np.float32))
        out = ops.ComputeLocalTsdf_hessian_batch(dev[set_frames[f]], intr, (res,) * 3, 7.68 / res, pb, trunc, gt)
        for (i, _), row in zip(mine_f, out):
            rows_local[i] = row
    torch.cuda.synchronize()
    t_loss = time.perf_counter() - t0
    if comm is not None:
        send = torch.from_numpy(rows_local.astype(np.float32).reshape(-1)).cuda()
        recv = torch.zeros((world * send.numel(),), dtype=torch.float32, device="cuda")
        xs._capi.check(xs.load().xs_comm_all_gather(comm.h, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()), send.numel(), None), "all_gather")
        torch.cuda.synchronize()
        g = recv.cpu().numpy().reshape(world, -1, 4).astype(np.float64)
        rows = np.stack([g[d % world][d // world] for d in range(len(all_dirs))])
    else:
        rows = rows_local
    if rank == 0:
        h = float(xs.H_)
        part_b = {"pose_set_frames": set_frames, "directions": len(all_dirs), "directions_rank0": len(my_dirs), "seconds_rank0": t_loss, "per_frame": []}
        ok = True
        for f in range(nset):
            r = rows[f * len(pairs):(f + 1) * len(pairs)]
            grad_by_pair = {}
            Hm = np.zeros((6, 6))
            for p, (i, j) in enumerate(pairs):
                Hm[i, j] = Hm[j, i] = r[p, 2] / h / h
                # the loss sums raw components: the eps1 part is the gradient along e_i (the reference sums real().imag())
                grad_by_pair.setdefault(i, []).append(r[p, 1] / h)
            grad = np.array([np.mean(grad_by_pair[i]) for i in range(6)])
            spread = max(float(np.ptp(grad_by_pair[i]) / max(abs(grad[i]), 1e-30)) for i in range(6))
            ev = np.linalg.eigvalsh(Hm)
            part_b["per_frame"].append({"loss": float(r[0, 0]), "voxels": float(r[0, 3]), "gradient": grad.tolist(), "gradient_spread_over_pairs_rel": spread,
                                        "hessian_diag": np.diag(Hm).tolist(), "hessian_min_eig": float(ev[0]), "hessian_max_eig": float(ev[-1])})
            ok = ok and np.isfinite(r).all() and r[0, 3] > 1000 and spread < 1e-3 and bool((r[:, 3] == r[0, 3]).all())
        # one row against the single-direction entry point (xs_tsdf_hessian): bit-identical by construction of the batch sweep
        f, p = all_dirs[0]
        i, j = pairs[p]
        xi = np.zeros((4, 6), np.float32)
        xi[1, i] = xs.H_
        xi[2, j] = xs.H_
        T = xs.se3_exp(xi, comps=3, dirs=1).astype(np.float64)
        v2c0 = np.linalg.inv(w2v @ xs.synth_pose(set_frames[f]).astype(np.float64))
        M = [T[c] @ v2c0 for c in range(1, 4)]
        pb = ops.PoseBatch(v2c0[:3, :3], v2c0[:3, 3], np.asarray([m[:3, :3].reshape(9) for m in M], np.float32), np.asarray([m[:3, 3] for m in M], np.float32))
        single = np.asarray(ops.ComputeLocalTsdf_hessian(dev[set_frames[f]], intr, (res,) * 3, 7.68 / res, pb, trunc, gt))
        part_b["row0_vs_single_call_rel"] = float(np.abs(single.astype(np.float32) - rows[0].astype(np.float32)).max() / max(np.abs(single).max(), 1e-30))
        ok = ok and part_b["row0_vs_single_call_rel"] <= 1e-6
        line = {"config": "configs[4]: %d^3 TSDF, %d first-order directions through the frame loop + relocalisation Hessian over a %d-frame pose set (%d bicomplex directions)"
                          % (res, ndirs, nset, len(all_dirs)), "n_gpus": world, "frame_loop": part_a, "relocalisation_hessian": part_b, "ok": bool(ok)}
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "config5_n%d.json" % world), "w") as fh:
            json.dump(line, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
