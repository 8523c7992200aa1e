"""BASELINE.json config 4: ExpectedImprovement over 4 M candidate points, N = 16384 2-D SquaredExponential, then 64
L-BFGS-B restarts of -ln EI from the best candidates.  One process per GPU (torchrun): candidates and restarts are
sharded over ranks, the only exchanges are an all_gather of each rank's top-k (value, point) pairs and of the
restart optima (torch.distributed, a few KB).  Parity: EI of a sampled subset against the oracle (--check)."""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import inference_tools_b200.gp as gp
from inference_tools_b200.sharding import round_robin, shard_range
from inference_tools_b200.gp._lockstep import lockstep_lbfgs


def synth(seed, n, d, sigma_n=0.05):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, sigma_n, n)
    return x, y, np.full(n, sigma_n)

from scipy.optimize import fmin_l_bfgs_b

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16384)
ap.add_argument("--cands", type=int, default=1 << 22)
ap.add_argument("--restarts", type=int, default=64)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = 2
x, y, e = synth(11, a.size, d, sigma_n=0.1)
theta = np.array([0.3, 0.1, np.log(0.3), np.log(0.3)])
t0 = time.perf_counter()
m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta, device=local)
fit_s = time.perf_counter() - t0
ei = gp.ExpectedImprovement(); ei.update_gp(m)
lo, hi = shard_range(a.cands, rank, world)
cand = np.random.default_rng(1234).uniform(0, 1, (a.cands, d))[lo:hi]
def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
ei.opt_func_batch(cand[:1024])          # warm-up
sync(); t0 = time.perf_counter()
score = ei.opt_func_batch(cand)         # -ln EI for this rank's slab
sync(); sweep_s = time.perf_counter() - t0
k = min(a.restarts, score.size)
top = np.argsort(score)[:k]
pairs = np.concatenate([score[top, None], cand[top]], axis=1)
if world > 1:
    buf = [None] * world
    dist.all_gather_object(buf, pairs)
    pairs = np.concatenate(buf)
pairs = pairs[np.argsort(pairs[:, 0])][: a.restarts]
mine = round_robin(len(pairs), rank, world)
bounds = [(0.0, 1.0)] * d
sync(); t0 = time.perf_counter()
if os.environ.get("GPB_CFG4_SEQUENTIAL"):   # the reference's way: one single-point prediction per evaluation
    res = [fmin_l_bfgs_b(ei.opt_func_gradient, pairs[i, 1:], approx_grad=False, bounds=bounds, pgtol=1e-10, maxiter=30) for i in mine]
else:                                        # the same scipy runs in lockstep, one batched device call per round
    res = lockstep_lbfgs(ei.opt_func_gradient_batch, [pairs[i, 1:] for i in mine], bounds, pgtol=1e-10, maxiter=30)
sync(); restart_s = time.perf_counter() - t0
best = min(((float(r[1]), r[0].tolist()) for r in res), default=(np.inf, None))
if world > 1:
    buf = [None] * world
    dist.all_gather_object(buf, best)
    best = min(buf)
tt = torch.tensor([sweep_s, restart_s], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    npad = (a.size + 127) // 128 * 128
    out = {"config": f"cfg4: EI over {a.cands} candidates, N={a.size} 2D SE, {a.restarts} restarts, {world} GPU(s)",
           "fit_s": fit_s, "sweep_s": tt[0].item(), "restarts_s": tt[1].item(), "candidates_per_s": a.cands / tt[0].item(),
           "sweep_tflops_aggregate": a.cands * float(npad) ** 2 / tt[0].item() / 1e12, "best_neg_log_ei": best[0], "best_x": best[1],
           "n_function_evals_rank0": int(sum(r[2]["funcalls"] for r in res)),
           "restart_mode": "sequential" if os.environ.get("GPB_CFG4_SEQUENTIAL") else "lockstep"}
    if a.check:
        from oracle import gp_oracle as orc
        f = orc.Fit(x, y, ("SE",), "const", theta, e**2)
        sub = cand[:: max(1, len(cand) // 256)][:256]
        mu, sig = f.predict(sub)
        ref = orc.neg_log_ei(mu, sig, y.max())
        got = ei.opt_func_batch(sub)
        out["parity_neg_log_ei_rel"] = float(np.abs(got - ref).max() / np.abs(ref).max())
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/cfg4_N{a.size}_g{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
