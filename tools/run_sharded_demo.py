"""run(estimator, tol) sharded over the GPUs of one node (torchrun, one rank per GPU, NCCL all-gather of the moment triplets)
against the same run on rank 0's GPU alone: identical adaptive decisions, estimates equal to 1e-12.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/run_sharded_demo.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import mle_b200
import mle_b200.host as H
import mle_b200.kl as kl

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = mle_b200.Engine(local)
m, levels = 100, 6
spec = H.ModelSpec("lognormal1d", 1, [16], {l: kl.kl_basis_1d(l, m=m) for l in range(levels + 1)})


def make(backend):
    return H.Estimator(H.ML(), H.QMC(), spec, [H.Normal()] * m, backend, nb_of_shifts=16, max_index_set_param=levels,
                       cost_model=lambda i: 2.0 ** i[0], shift_seed=11, nb_of_tols=6)


tol = 2e-5
t0 = time.perf_counter()
hs = H.run(make(H.ShardedBackend(H.GpuBackend(eng), device=torch.device("cuda", local))), tol)
t_sharded = time.perf_counter() - t0
dist.barrier()
if rank == 0:
    t0 = time.perf_counter()
    h1 = H.run(make(H.GpuBackend(eng)), tol)
    t_single = time.perf_counter() - t0
    same_n = all(hs[i]["nb_of_samples"] == h1[i]["nb_of_samples"] for i in range(len(h1)))
    rel = max(abs(hs[i]["mean"] - h1[i]["mean"]) / abs(h1[i]["mean"]) for i in range(len(h1)))
    relv = max(abs(hs["dV"][k] - h1["dV"][k]) / abs(h1["dV"][k]) for k in h1["dV"] if h1["dV"][k] == h1["dV"][k] and h1["dV"][k] != 0)
    print(json.dumps({"world": world, "tol": tol, "levels": len(h1["current_index_set"]), "mean": hs["mean"], "rmse": hs["rmse"],
                      "nb_of_samples": {str(k): v for k, v in hs["nb_of_samples"].items() if v},
                      "same_sample_counts_as_single_gpu": same_n, "max_rel_diff_mean": rel, "max_rel_diff_level_variances": relv,
                      "seconds_sharded": t_sharded, "seconds_single_gpu": t_single}))
    assert same_n and rel <= 1e-12 and relv <= 1e-10
dist.barrier()
dist.destroy_process_group()
