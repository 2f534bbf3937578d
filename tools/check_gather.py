#!/usr/bin/env python
"""Hardware check of the multi-GPU layer (run under torchrun on N >= 2 GPUs of one box):
the N-rank run - every rank carries its block of the second-order pairs and the first-order components of the parameters
those pairs touch (parallel.plan_hessian_shards), the library
all-gathers the pose records over NCCL after every frame (csrc/comm.cpp) - against the 1-rank run of the full batch on rank 0:
the gathered record must equal the full record: bit for bit on the real part, <= 1e-6 relative on first-order and <= 3e-6 on
second-order components (a rank's share and the full batch sum their ICP normal equations in different orders).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_gather.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import xslam_b200 as xs
    from xslam_b200 import parallel
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = int(os.environ.get("XS_CHECK_RES", "256"))
    frames = int(os.environ.get("XS_CHECK_FRAMES", "4"))
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=res, tsdf_size_y=res, tsdf_size_z=res, tsdf_voxel_size=7.68 / res)
    rng = np.random.default_rng(7)
    U = np.concatenate([np.eye(6), rng.standard_normal((2, 6)) / np.sqrt(6)])  # 8 parameters, 36 pairs
    n = U.shape[0]
    pairs = xs.all_pairs(n)
    plan = parallel.plan_hessian_shards(n, pairs, world)  # blocked shards: a rank carries only the parameters its pairs touch
    mine = plan[rank]
    seeds, _ = xs.hessian_seeds(U[mine["params"]], mine["local_pairs"])
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=2, seeds=seeds, pairs=mine["local_pairs"], n_params=len(mine["params"]))
    comm = parallel.Comm.from_torch_distributed(dist, device="cuda")
    L = parallel.planned_record_floats(plan)
    k.set_comm(comm, L)
    k.set_deferred(True)
    full, g_prev = None, None
    if rank == 0:
        sf, _ = xs.hessian_seeds(U, pairs)
        full = xs.KinectFusionReconstruction()
        full.SetYamlParameters(cfg, comps=2, seeds=sf, pairs=pairs, n_params=n)
    rep = {"world": world, "parameters": n, "pairs": len(pairs), "planes_per_rank": [len(sh["params"]) + len(sh["pair_ids"]) for sh in plan],
           "frames": []}
    ok = True
    for f in range(frames):
        d = xs.synth_depth(f)
        assert k.ProcessFrame(d) == 1
        g = k.gathered_records()
        if f > 0:  # the two gather buffers alternate: the previous frame's records stay readable while this frame's gather runs
            ok = ok and bool(np.array_equal(k.gathered_records(lag=1), g_prev))
        g_prev = np.array(g, copy=True)
        g = np.asarray(g.cpu() if hasattr(g, "cpu") else g)
        rec = parallel.assemble_planned_records(g, plan, n, len(pairs))
        # replicas: every rank produced the same real pose, and the ranks that share a parameter the same first-order component
        gr = g.reshape(world, -1, 16)
        same_real = bool((gr[:, 0] == gr[0, 0]).all())
        # (the ranks hold different component sets, so their ICP passes sum in different orders and may even take different
        # kernel forms: shared first-order components agree to FP32 summation noise, not bit for bit)
        sc = max(np.abs(rec[1:1 + n]).max(), 1e-30)
        first_spread = max(float(np.abs(gr[r][1 + i] - rec[1 + q]).max() / sc) for r, sh in enumerate(plan) for i, q in enumerate(sh["params"]))
        same_first = first_spread <= 1e-6
        fr = {"frame": f, "replicas_real_identical": same_real, "replicas_first_order_spread_rel": first_spread}
        if rank == 0:
            assert full.ProcessFrame(d) == 1
            w = full.world2camera.reshape(-1, 16)
            fr["real_identical_to_1rank"] = bool(np.array_equal(rec[0], w[0]))
            sc1 = max(np.abs(w[1:1 + n]).max(), 1e-30)
            sc2 = max(np.abs(w[1 + n:]).max(), 1e-30)
            fr["first_order_rel"] = float(np.abs(rec[1:1 + n] - w[1:1 + n]).max() / sc1)
            fr["second_order_rel"] = float(np.abs(rec[1 + n:] - w[1 + n:]).max() / sc2)
            ok = ok and fr["real_identical_to_1rank"] and fr["first_order_rel"] <= 1e-6 and fr["second_order_rel"] <= 3e-6
        ok = ok and same_real and same_first
        rep["frames"].append(fr)
    k.sync()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        rep["ok"] = bool(flag.item())
        print(json.dumps(rep), flush=True)
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "gather_check_n%d.json" % world), "w") as fh:
            json.dump(rep, fh, indent=1)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
