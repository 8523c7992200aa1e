"""Benchmark of the GpRegressor hot path on B200 (contract: see the round prompt / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload = BASELINE.json config 3, the shape the headline metric is quoted on: GpRegressor with
RationalQuadratic + WhiteNoise in 5-D, N = 32768 training points, fixed well-conditioned theta.  One step =
one marginal_likelihood_gradient(theta) (assemble, Cholesky, explicit inverse, traces) + one
set_hyperparameters(theta) (assemble, Cholesky, alpha) + predict mean/sigma at 131072 query points per GPU
(2^20 over 8 GPUs: weak scaling, query points are the sharded unit, no data-path collective).
`value` = query points processed by all ranks / max-over-ranks device time of the step (fit and gradient
time included), inputs resident in HBM; `e2e` = the same through the public GpRegressor API with pinned
host buffers.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN = int(os.environ.get("GPB_BENCH_N", 32768))
DIM = 5
M_PER_GPU = int(os.environ.get("GPB_BENCH_M", 131072))
COMPS = ("RQ", "WHITE")
THETA = np.array([0.2, 0.1, 1.0] + [np.log(0.3)] * DIM + [np.log(0.05)])
METRIC = "GpRegressor fit+gradient+predict throughput at N=32768,d=5 (query points/s over the whole step; seconds in step_s)"
UNIT = "query points/s"
SAMPLE_N = int(os.environ.get("GPB_BENCH_CPU_N", 8192))       # --impl reference: largest sample of the unmodified reference
SAMPLE_N_SMALL = int(os.environ.get("GPB_BENCH_CPU_N_SMALL", 4096))  # second size for the N^2 / N^3 split; cpu_baseline leg
SAMPLE_M = 32


def synth(seed, n, d, sigma_n=0.05):
    """Seeded synthetic regression problem of SURVEY.md section 8d (the oracle keeps its own copy)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, sigma_n, n)
    return x, y, np.full(n, sigma_n)


def workload_config(n_gpus):
    return {
        "workload": "cfg3: GpRegressor RationalQuadratic+WhiteNoise 5D, N=32768; step = marginal_likelihood_gradient + "
                    "set_hyperparameters + predict mean/sigma at 131072 points per GPU",
        "N": N_TRAIN, "d": DIM, "points_per_gpu": M_PER_GPU, "points_total": M_PER_GPU * n_gpus,
        "theta": "a=e^0.1, alpha=e, l=0.3, white=0.05, y_err=0.05 (SURVEY.md 8d)",
        "l2": "inputs larger than L2: K/L are 8.6 GB each, query slab 5.2 MB re-read against 8.6 GB of L",
        "sharding": f"query points, {n_gpus} rank(s), fit replicated per rank, no collective on the data path",
    }


def load_fp64_peak():
    """FP64 tensor (DMMA) peak: MEASURED_PEAKS.json carries only HBM and bf16; the FP64 denominator is this
    repo's own measurement on the same pool (tools/fp64_probe.cu -> profiles/fp64_peak_r1.json)."""
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r1.json")))
        return float(p["probe"]["dmma_tflops_sustained"]), "measured: tools/fp64_probe DMMA issue rate (profiles/fp64_peak_r1.json); MEASURED_PEAKS.json has no FP64 entry"
    except Exception:
        return 37.2, "nominal 148 SM x 128 flop/clk x 1.965 GHz (profiles/fp64_peak_r1.json missing)"


class ClockSampler:
    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0]))
            smax = int(f[1])
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port(n_gpus, n_sample):
    """Fallback when no copy of the reference is on this host: the oracle port of its algorithm (same LAPACK/BLAS calls)."""
    from oracle import cpu_reference as cr
    t = cr.timed_step(n_sample, DIM, COMPS, "const", THETA, SAMPLE_M)
    total_pts = M_PER_GPU * n_gpus
    full_s, fit_s, pred_s = cr.extrapolate(t, n_sample, N_TRAIN, total_pts)
    return {
        "value": total_pts / full_s, "unit": UNIT, "cores": cr.blas_threads(), "kind": "port",
        "sample": f"reference algorithm (oracle port: numpy/scipy LAPACK+BLAS, the reference's own calls) timed at N={n_sample}, "
                  f"{SAMPLE_M} query points; components extrapolated to N={N_TRAIN}, M={total_pts}: assembly/traces x(N/Ns)^2, "
                  f"LAPACK x(N/Ns)^3, predict per point x(N/Ns)^2 x M. No copy of the unmodified reference on this host.",
        "sample_seconds": {k: round(v, 4) for k, v in t.items() if not k.startswith("_")},
        "extrapolated_step_s": full_s, "extrapolated_fit_grad_s": fit_s, "extrapolated_predict_s": pred_s,
    }


def cpu_reference(n_gpus, sizes):
    """The UNMODIFIED reference (baseline/_ref: inference.gp.GpRegressor with RationalQuadratic + WhiteNoise) timed on the
    host cores at the given sample sizes -- the full configuration needs two (N,N,d) arrays of 43 GB each plus p dense
    gradient planes and one Python-level dtrtrs per query point -- and extrapolated per call to the benchmark's N and M."""
    from oracle import cpu_reference as cr
    ts = [cr.reference_timed_step(n, DIM, THETA, SAMPLE_M) for n in sizes]
    if ts[0] is None:
        return cpu_port(n_gpus, sizes[-1])
    total_pts = M_PER_GPU * n_gpus
    if len(sizes) >= 2:
        ex = cr.reference_extrapolate(ts[0], sizes[0], ts[-1], sizes[-1], N_TRAIN, total_pts)
        how = f"two sizes N={sizes[0]} and N={sizes[-1]}: each call fitted as a N^2 + b N^3 and evaluated at N={N_TRAIN}"
    else:
        n1, r = sizes[0], N_TRAIN / sizes[0]
        ex = {"grad": ts[0]["grad"] * r**3, "fit": ts[0]["fit"] * r**3, "predict": ts[0]["predict_per_point"] * r**2 * total_pts}
        ex["step"] = ex["grad"] + ex["fit"] + ex["predict"]
        how = f"one size N={n1}: marginal_likelihood_gradient and set_hyperparameters scaled by (N/Ns)^3"
    return {
        "value": total_pts / ex["step"], "unit": UNIT, "cores": cr.blas_threads(), "kind": "reference",
        "sample": f"unmodified reference GpRegressor (baseline/_ref), RQ+White d={DIM}: marginal_likelihood_gradient + set_hyperparameters "
                  f"+ __call__ at {SAMPLE_M} points, {how}; predict per point x(N/Ns)^2 x M={total_pts}. The reference cannot "
                  f"allocate N={N_TRAIN}, d={DIM}.",
        "sample_seconds": {f"N={n}": {k: round(v, 4) for k, v in t.items() if not k.startswith("_")} for n, t in zip(sizes, ts)},
        "extrapolated_step_s": ex["step"], "extrapolated_fit_grad_s": ex["grad"] + ex["fit"], "extrapolated_predict_s": ex["predict"],
    }


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step on this host's cores.  One "step" here is
    one timed pass of the unmodified reference at the two sample sizes (about a minute); the run is capped at 2 steps and
    no warm-up so that it ends within a few minutes (the first import / BLAS thread start-up is inside the first pass)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = 0
    steps = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    vals = [cpu_reference(args.gpus, [SAMPLE_N_SMALL, SAMPLE_N]) for _ in range(steps)]
    wall = time.perf_counter() - t0
    best = max(vals, key=lambda v: v["value"])
    value = float(np.mean([v["value"] for v in vals]))
    best = dict(best, value=value)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * args.gpus * M_PER_GPU / value, "sample_wall_s_per_step": wall / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": best, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def dist_cholesky_block(args, rank, world, local, gp_engine, barrier, max_over_ranks):
    """BASELINE config 5 beside the headline number: GpRegressor SquaredExponential 2-D, N = 131072, block-column-cyclic
    FP64 Cholesky + log marginal likelihood over the `world` ranks of this run (csrc/dist.cu: NCCL panel broadcasts, INT8
    tensor-core trailing updates and panel solves).  Reported: seconds (max over ranks, CUDA events), aggregate and
    per-GPU FP64-equivalent TFLOP/s (N^3/3), the LML (must agree across world sizes) and the strong-scaling efficiency
    against the one-GPU sweep committed under profiles/.  GPB_BENCH_DIST_N=0 skips it."""
    n = int(os.environ.get("GPB_BENCH_DIST_N", 131072))
    if n <= 0:
        return None
    import torch.distributed as dist
    from inference_tools_b200 import _lib
    gp_engine.close()                      # release the headline workload's HBM before the 64 GiB factor
    d, block = 2, int(os.environ.get("GPB_BENCH_DIST_BLOCK", 2048))
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, 0.05, n)
    theta = np.array([0.2, 0.1] + [np.log(0.3)] * d)
    uid = [_lib.nccl_unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    eng = _lib.Engine(local)
    eng.set_data(x, y, np.full(n, 0.05**2))
    eng.set_model([_lib.COV_SE], _lib.MEAN_CONST)
    eng.dist_init(rank, world, uid[0])
    runs = []
    for _ in range(2):                     # one warm-up sweep (allocations, kernel attributes, NCCL channels), one timed
        barrier()
        lml, info, t = eng.dist_lml(theta, block)
        runs.append((lml, info, max_over_ranks(t["factor_s"]), max_over_ranks(t["assemble_s"])))
    grad_block = None
    if world >= 4 and os.environ.get("GPB_BENCH_DIST_GRAD", "1") != "0":
        # the gradient of the same likelihood on the same layout (gpb_dist_lml_grad): N^2 / world doubles of inverse rows per
        # rank, so only from 4 ranks up at this N.  Reported, never part of the headline value; a failure is reported too.
        ok, secs, err, res = 1.0, (0.0, 0.0), "", None
        barrier()
        try:
            res = eng.dist_lml_grad(theta, block)
            secs = (res[3]["factor_s"], res[3]["gradient_s"])
        except Exception as exc:  # noqa: BLE001 -- the bench line must still be printed
            ok, err = 0.0, str(exc)[:300]
        # every rank takes part in the same three reductions whether its call failed or not
        ok_all = -max_over_ranks(-ok)
        factor_g, gradient_g = max_over_ranks(secs[0]), max_over_ranks(secs[1])
        if ok_all > 0.5:
            grad_block = {"lml": res[0], "info": res[2], "grad": [float(v) for v in res[1]], "factor_seconds": factor_g,
                          "gradient_seconds": gradient_g,
                          "what": "alpha + rows of K^-1 (streamed solves, Y_a Y_b^T products) + traces + all-reduce, after the factor"}
        else:
            grad_block = {"error": err or "failed on another rank"}
    eng.dist_finalize()
    eng.close()
    lml, info, factor_s, assemble_s = runs[-1]
    npad = (n + 127) // 128 * 128
    out = {"workload": f"cfg5: SquaredExponential 2D, N={n}, block-column-cyclic Cholesky + LML, block {block}", "world": world,
           "factor_seconds": factor_s, "assemble_seconds": assemble_s, "info": info, "lml": lml,
           "fp64_equiv_tflops_aggregate": npad**3 / 3 / factor_s / 1e12, "fp64_equiv_tflops_per_gpu": npad**3 / 3 / factor_s / 1e12 / world,
           "collective": "ncclBroadcast per panel (N^2/2 x 8 B received per rank in total) + one 2-double ncclAllReduce"}
    if grad_block is not None:
        out["gradient"] = grad_block
    try:
        ref = json.load(open(os.path.join(ROOT, "profiles", "dist_cholesky_1gpu_ref_r2.json")))
        if ref.get("n") == n and ref.get("block") == block:
            out["one_gpu_factor_seconds_ref"] = ref["factor_s"]
            out["efficiency_vs_one_gpu_ref"] = ref["factor_s"] / (world * factor_s)
            out["lml_rel_diff_vs_one_gpu_ref"] = abs(lml - ref["lml"]) / abs(ref["lml"])
    except Exception:
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints a version banner on stdout when a
    # communicator is created): everything but the final line is sent to stderr at the file-descriptor level.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    import torch
    import torch.distributed as dist
    from inference_tools_b200 import _lib
    from inference_tools_b200.gp import GpRegressor, RationalQuadratic, WhiteNoise
    from inference_tools_b200.sharding import shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- synthetic data (identical on every rank), query slab of this rank
    x, y, y_err = synth(2024, N_TRAIN, DIM)
    m_total = M_PER_GPU * world
    lo, hi = shard_range(m_total, rank, world)
    q_all_rng = np.random.default_rng(99)
    q_host = torch.empty((hi - lo, DIM), dtype=torch.float64, pin_memory=True)
    q_np = q_host.numpy()
    q_np[:] = q_all_rng.uniform(0, 1, (m_total, DIM))[lo:hi]

    gp = GpRegressor(x, y, y_err=y_err, kernel=RationalQuadratic() + WhiteNoise(), hyperpars=THETA, device=local)
    eng = gp.engine
    m_loc = hi - lo
    q_dev = eng.dev_alloc(m_loc * DIM)
    mu_dev = eng.dev_alloc(m_loc)
    sig_dev = eng.dev_alloc(m_loc)
    eng.dev_upload(q_dev, q_np)

    phase_acc = {}

    def add_phases(prefix, tm):
        for k, v in tm.items():
            phase_acc[prefix + k] = phase_acc.get(prefix + k, 0.0) + v

    def step_resident(record=False):
        """device-resident step; returns device milliseconds from the library's CUDA events"""
        ms = 0.0
        lml, grad, info = eng.lml_grad(THETA)
        tm = eng.timers(); ms += sum(tm.values())
        if record: add_phases("grad.", tm)
        info2 = eng.factor(THETA)
        tm = eng.timers(); ms += sum(tm.values())
        if record: add_phases("fit.", tm)
        eng.predict_dev(q_dev, m_loc, mu_dev, sig_dev)
        eng.sync()
        tm = eng.timers(); ms += sum(tm.values())
        if record: add_phases("predict.", tm)
        assert info == 0 and info2 == 0
        return ms, lml, grad

    def step_e2e():
        lml, grad = gp.marginal_likelihood_gradient(THETA)
        gp.set_hyperparameters(THETA)
        mu, sig = gp(q_np)
        return float(lml), mu, sig

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0, flops0, flops_i8_0 = eng.launch_count(), eng.gemm_flops(), eng.gemm_flops_int8()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        ms, lml, grad = step_resident(record=True)
        dev_ms += ms
    barrier()
    wall_s = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    flops = eng.gemm_flops() - flops0
    flops_i8 = eng.gemm_flops_int8() - flops_i8_0
    clocks = sampler.stop() if sampler else None
    dev_s = max_over_ranks(dev_ms * 1e-3)
    wall_s = max_over_ranks(wall_s)
    step_s = dev_s / args.steps

    # ---- e2e through the public API with pinned host buffers
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lml_e, mu_h, sig_h = step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps

    # parity guard inside the bench: resident and e2e paths agree bit for bit on this rank's slab
    mu_r = eng.dev_download(mu_dev, m_loc)
    assert np.array_equal(mu_r, mu_h), "device-resident and host-buffer paths disagree"

    dist_block = dist_cholesky_block(args, rank, world, local, gp_engine=eng, barrier=barrier, max_over_ranks=max_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_fp64_peak()
    ph = {k: v / args.steps for k, v in phase_acc.items()}
    potrf_ms = 0.5 * (ph.get("grad.potrf", 0) + ph.get("fit.potrf", 0))
    chol_tf = (N_TRAIN**3 / 3) / (potrf_ms * 1e-3) / 1e12 if potrf_ms else None
    npad = (N_TRAIN + 127) // 128 * 128
    # dominant kernel = gemm_i8_kernel: FP64 GEMMs executed as 28 exact int8 tensor-core products (csrc/gemm_i8.cu).
    # Roofline: int8 operations executed on the tcgen05 pipe / CUDA-event time of the GEMM-only phases, against twice
    # the MEASURED sustained bf16 rate (kind::i8 issues at twice the kind::f16 rate; MEASURED_PEAKS.json has no int8 entry).
    gemm_phase_ms = sum(ph.get(k, 0) for k in ("grad.potrf", "fit.potrf", "grad.trtri", "grad.lauum", "predict.planes", "predict.trsm"))
    gemm_tf = (flops / args.steps) / (gemm_phase_ms * 1e-3) / 1e12 if gemm_phase_ms else None
    int8_tops = 28.0 * (flops_i8 / args.steps) / (gemm_phase_ms * 1e-3) / 1e12 if gemm_phase_ms else None
    pred_tf = (m_loc * float(npad) ** 2) / (ph.get("predict.trsm", 1e30) * 1e-3) / 1e12
    # Denominator: the MEASURED sustained full-chip rate of tcgen05.mma kind::i8 (tools/int8_peak.cu: every SM issuing
    # back-to-back 128x128x32 MMAs on random digits for 4 s; the chip runs into its power limit there exactly as in this
    # step, ~1.6 GHz).  MEASURED_PEAKS.json has no int8 entry; without the probe file fall back to 2 x its sustained bf16.
    int8_peak, int8_src = 2 * 1397.2, "fallback: 2 x 1397.2 (bf16 sustained of this pool when MEASURED_PEAKS.json was written)"
    try:
        pk = json.load(open(os.path.join(ROOT, "profiles", "int8_peak_r2.json")))
        int8_peak = float(pk["sustained_int8_tops"])
        int8_src = (f"measured: tools/int8_peak.cu sustained {pk['sustained_seconds']:.1f} s on all SMs, SM clock {pk['clocks']['sm_mhz']} MHz "
                    f"under {pk['clocks']['reasons']} (profiles/int8_peak_r2.json); MEASURED_PEAKS.json has no int8 entry "
                    f"(2 x its sustained bf16 would be {2 * 1397.2:.0f})")
    except Exception:
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            int8_peak = 2.0 * float(mp["bf16_tflops_sustained"])
            int8_src = "2 x MEASURED_PEAKS.json bf16_tflops_sustained (profiles/int8_peak_r2.json missing)"
        except Exception:
            pass
    traffic = None
    try:
        for name in ("gemm_i8_traffic_r2.json", "gemm_i8_traffic_r1.json"):
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                traffic = json.load(open(path)).get("dram_bytes_per_launch")
                break
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": m_total / step_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "step_s": step_s, "wall_ms_per_step": wall_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world), "clocks": clocks,
        "e2e": {"value": m_total / e2e_s, "unit": UNIT, "step_s": e2e_s,
                "h2d_bytes_per_step": int(m_loc * DIM * 8 + 3 * THETA.size * 8), "d2h_bytes_per_step": int(2 * m_loc * 8 + (THETA.size + 1) * 8)},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "gemm_i8_kernel (tcgen05.mma kind::i8 -> UTCIMMA; 28 exact int8 products per FP64 product)",
                     "bound": "tensor", "achieved": int8_tops, "peak": int8_peak, "unit": "TFLOP/s",
                     "frac": (int8_tops / int8_peak) if int8_tops else None, "traffic": traffic, "peak_source": int8_src,
                     "how": "28 x algorithmic FP64 GEMM flops issued on the INT8 path in the step / CUDA-event time of the GEMM-only phases (potrf x2, trtri, lauum, predict planes + solve; these also hold the operand splitting, the short-k DMMA GEMMs and the diagonal-block kernels) on the library stream; traffic = ncu dram bytes of one launch of the predict shape",
                     "fp64_equivalent": {"achieved": gemm_tf, "fp64_dmma_peak": peak, "ratio": (gemm_tf / peak) if gemm_tf else None,
                                         "int8_share_of_gemm_flops": (flops_i8 / flops) if flops else None, "fp64_peak_source": peak_src}},
        "cholesky": {"seconds": potrf_ms * 1e-3, "tflops": chol_tf, "frac_of_fp64_peak": chol_tf / peak if chol_tf else None, "flops": "N^3/3"},
        "predict": {"tflops": pred_tf, "frac_of_fp64_peak": pred_tf / peak, "flops": "M N^2"},
        "phases_ms": {k: round(v, 3) for k, v in sorted(ph.items())},
        "lml": float(lml),
    }
    line["dist_cholesky"] = dist_block
    if world == 1:
        line["cpu_baseline"] = cpu_reference(1, [SAMPLE_N_SMALL])   # ~10-20 s of CPU work on the unmodified reference
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
