"""BASELINE config 5: a batch of independent start/goal problems on one shared map, sharded over the GPUs of one box.
Launch with torchrun (one rank per GPU) or plain python (1 GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/run_batch.py --problems 4096

Flow: rank 0 builds the occupancy map and packs it (generateMapKernel2D layout, PCSmap_manager.h:81-108) -> NCCL broadcast
of the packed bytes (the ONLY collective before the run) -> every rank opens a POOL of contexts on its GPU (own stream each)
that adopt the broadcast device buffer (svsdf_set_map_device) -> problems are drawn from ONE queue shared by all ranks
(atomic counter in the torch.distributed store, longest problem first) by the pool's worker threads (svsdf_optimize_batch):
per problem the query points are built ON THE DEVICE from the map (svsdf_extract_points) and the trajectory is optimised
(svsdf_optimize: host MINCO + L-BFGS, cost and gradient on the GPU) -> a final all-reduce gathers the per-problem results.
Several contexts per GPU overlap one problem's host work and latency-bound kernel tails with another problem's kernels.
Prints one JSON line (rank 0): problems/s and aggregate query points/s (points x evaluations per second).

Problem set: first 1000 start/goal pairs from the reference's src/coords.txt (fixture copy under
tests/golden/), the rest uniform; star, 8-piece MINCO, ~200k query points each from the shared map; `--first K` runs the first K
problems of that set (weak-scaling runs use K proportional to the number of GPUs).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from implicit_svsdf_planner_b200 import api, batch, scenes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=4096, help="size of the problem set (config 5: 4096)")
    ap.add_argument("--first", type=int, default=0, help="solve only the first K problems of the set (0 = all)")
    ap.add_argument("--max-iter", type=int, default=0, help="L-BFGS iteration cap (0 = to termination)")
    ap.add_argument("--ctx-per-gpu", type=int, default=3)
    ap.add_argument("--res", type=float, default=0.025)
    ap.add_argument("--density", type=float, default=0.36, help="occupancy of the random map (0.36 -> ~200k points per problem)")
    ap.add_argument("--pieces", type=int, default=8)
    ap.add_argument("--static", action="store_true", help="contiguous static split instead of the shared queue")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        batch.bind_to_gpu_numa(local)  # host workers and staging memory next to this rank's GPU
        dist.init_process_group("nccl", device_id=dev)
    extent, ks = 60.0, 17
    n = int(np.ceil(extent / args.res))
    kern = None
    if rank == 0:
        gm = batch.make_random_map(extent=extent, res=args.res, density=args.density, seed=scenes.SEED_MAP)
        kern = batch.pack_map_kernel(gm.occ, ks)
    t_b0 = time.perf_counter()
    kt = batch.broadcast_map(kern, device=dev)  # uint8 tensor on this rank's GPU
    torch.cuda.synchronize()
    t_bcast = time.perf_counter() - t_b0
    ctxs = []
    for _ in range(max(1, args.ctx_per_gpu)):
        c = api.Context("star", device=local)
        c.set_map_device(kt.data_ptr(), n, n, ks, (0.0, 0.0), args.res)
        ctxs.append(c)

    problems = scenes.make_batch_problems(args.problems, seed=scenes.SEED_BATCH, extent=(8.0, 52.0))
    K = args.first if args.first > 0 else args.problems
    problems = problems[:K]
    half = scenes.YAML["kernel_size"] * scenes.YAML["occupancy_resolution"] / 3.0
    params = api.default_lbfgs_params(mem_size=16, past=3, delta=1e-6, g_epsilon=0.0, max_iterations=args.max_iter, min_step=1e-32)
    # host-side problem descriptions (start/goal -> seeded nominal spline, waypoints, keep-out corridor): cheap, done up front
    descr = []
    for k in range(K):
        sg = problems[k]
        init_s, final_s, q, T = scenes.make_trajectory("star", args.pieces, scenes.SEED_BATCH + k, sg[:2], sg[2:4])
        b = scenes.minco_dense(init_s, final_s, q, T)
        wps = np.concatenate([init_s[:2, :1], q[:2], final_s[:2, :1]], axis=1).T
        descr.append(dict(init_s=init_s, final_s=final_s, x0=np.concatenate([scenes.backward_T(T), q.T.reshape(-1)]),
                          waypoints=wps, half=half, keepout=batch.keepout_samples(b, T), clearance=2.75))
    # warm-up: one short problem per context (module load, first-launch costs, occupancy queries) — not timed
    wp = api.default_lbfgs_params(mem_size=16, past=3, delta=1e-6, g_epsilon=0.0, max_iterations=2, min_step=1e-32)
    api.optimize_batch(ctxs, [dict(descr[0]) for _ in ctxs], args.pieces, wp)

    if args.static:
        mine = list(batch.partition(K, world, rank))
        it = iter(mine)
        import threading
        lock = threading.Lock()

        def next_index():
            with lock:
                return next(it, -1)
        taken = mine
    else:
        q = batch.WorkQueue(K, batch.lpt_order(problems))
        next_index = q.next
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rc, X, st, status, npts = api.optimize_batch(ctxs, descr, args.pieces, params, next_index=next_index)
    torch.cuda.synchronize()
    my_s = time.perf_counter() - t0
    if not args.static:
        taken = list(q.taken)
    if rc != 0:
        print(f"rank {rank}: svsdf_optimize_batch failed with {rc}: {ctxs[0].last_error() if hasattr(ctxs[0], 'last_error') else ''}", file=sys.stderr)
        sys.exit(3)
    res = np.zeros((K, 7))  # points, evals, iters, final cost, gpu seconds, status, solver seconds
    for k in taken:
        res[k] = [npts[k], st[k]["evaluations"], st[k]["iterations"], st[k]["final_cost"], st[k]["gpu_seconds"], status[k], st[k]["seconds"]]
    R = batch.gather_rows(res, taken, device=dev)
    t = torch.tensor([my_s, float(len(taken))], dtype=torch.float64, device=dev)
    tl = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(tl, t)
    else:
        tl = [t]
    if rank == 0:
        per_rank_s = [float(v[0]) for v in tl]
        per_rank_n = [int(v[1]) for v in tl]
        wall = max(per_rank_s)
        ok = R[:, 5] >= 0
        pts_evals = float((R[:, 0] * R[:, 1]).sum())
        print(json.dumps({
            "config": "5: batch of start/goal problems (star, 8-piece MINCO, query points extracted on the device from one shared map), "
                      "full svsdf_optimize per problem" + ("" if args.max_iter == 0 else f" capped at {args.max_iter} iterations"),
            "n_gpus": world, "problem_set": args.problems, "problems_solved": K, "ctx_per_gpu": len(ctxs),
            "queue": "static contiguous split" if args.static else "shared dynamic queue (c10d store counter), LPT order",
            "mean_points_per_problem": float(R[:, 0].mean()), "total_evaluations": int(R[:, 1].sum()),
            "total_iterations": int(R[:, 2].sum()), "wall_seconds_max_over_ranks": wall, "wall_seconds_per_rank": per_rank_s,
            "problems_per_rank": per_rank_n, "problems_per_s": K / wall, "aggregate_query_pts_per_s": pts_evals / wall,
            "gpu_seconds_sum": float(R[:, 4].sum()), "gpu_busy_fraction": float(R[:, 4].sum()) / (wall * world),
            "status_nonnegative_fraction": float(ok.mean()), "status_counts": {str(int(v)): int((R[:, 5] == v).sum()) for v in np.unique(R[:, 5])},
            "map_bytes_broadcast": int(kt.numel()), "broadcast_seconds": t_bcast,
            "mean_final_cost": float(R[:, 3].mean()), "mean_evaluations_per_problem": float(R[:, 1].mean())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
