"""BASELINE config 5 (reduced count by default): a batch of independent start/goal problems on one shared map, sharded
over the GPUs of one box.  Launch with torchrun (one rank per GPU) or plain python (1 GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/run_batch.py --problems 64 --max-iter 30

Flow per rank: rank 0 builds the random occupancy map and packs it (generateMapKernel2D layout) -> NCCL broadcast of
the packed bytes (the only collective before the run) -> every rank adopts the broadcast device buffer
(svsdf_set_map_device) -> for each of its problems: query points are built ON THE DEVICE from the map (K3,
svsdf_extract_points) and the trajectory is optimised (svsdf_optimize) -> final all-reduce gathers per-problem results.
Prints one JSON line (rank 0)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from implicit_svsdf_planner_b200 import api, batch, scenes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=32)
    ap.add_argument("--max-iter", type=int, default=30)
    ap.add_argument("--res", type=float, default=0.025)
    ap.add_argument("--density", type=float, default=0.27)
    ap.add_argument("--pieces", type=int, default=8)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
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
    ctx = api.Context("star", device=local)
    ctx.set_map_device(kt.data_ptr(), n, n, ks, (0.0, 0.0), args.res)
    problems = scenes.make_batch_problems(args.problems, seed=scenes.SEED_BATCH, extent=(8.0, 52.0))
    mine = batch.partition(args.problems, world, rank)
    half = scenes.YAML["kernel_size"] * scenes.YAML["occupancy_resolution"] / 3.0
    params = api.default_lbfgs_params(mem_size=16, past=3, delta=1e-6, g_epsilon=0.0, max_iterations=args.max_iter, min_step=1e-32)
    res = np.zeros((args.problems, 6))  # points, evals, iters, final cost, gpu seconds, extract seconds
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in mine:
        sg = problems[k]
        init_s, final_s, q, T = scenes.make_trajectory("star", args.pieces, scenes.SEED_BATCH + k, sg[:2], sg[2:4])
        b = scenes.minco_dense(init_s, final_s, q, T)
        wps = np.concatenate([init_s[:2, :1], q[:2], final_s[:2, :1]], axis=1).T
        te = time.perf_counter()
        P = ctx.extract_points(wps, half, batch.keepout_samples(b, T), 2.75)  # K3, on the device
        te = time.perf_counter() - te
        x0 = np.concatenate([scenes.backward_T(T), q.T.reshape(-1)])
        rc, x, To, bo, st = ctx.optimize(init_s, final_s, x0, args.pieces, params)
        res[k] = [P, st["evaluations"], st["iterations"], st["final_cost"], st["gpu_seconds"], te]
    torch.cuda.synchronize()
    my_s = time.perf_counter() - t0
    t = torch.tensor([my_s], dtype=torch.float64, device=dev)
    r = torch.from_numpy(res).to(dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)  # rows are disjoint across ranks
    if rank == 0:
        R = r.cpu().numpy(); wall = float(t.item())
        pts_evals = float((R[:, 0] * R[:, 1]).sum())
        print(json.dumps({"config": "5 (reduced)", "n_gpus": world, "problems": args.problems, "max_iter": args.max_iter,
                          "mean_points_per_problem": float(R[:, 0].mean()), "total_evaluations": int(R[:, 1].sum()),
                          "total_iterations": int(R[:, 2].sum()), "wall_seconds_max_over_ranks": wall,
                          "problems_per_s": args.problems / wall, "aggregate_query_pts_per_s": pts_evals / wall,
                          "gpu_seconds_sum": float(R[:, 4].sum()), "extract_seconds_sum": float(R[:, 5].sum()),
                          "map_bytes_broadcast": int(kt.numel()), "broadcast_seconds": t_bcast,
                          "mean_final_cost": float(R[:, 3].mean())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
