#!/usr/bin/env python3
"""bench.py -- RM2 users scored/sec at ML-20M shape (BASELINE.json metric), 1/2/4/8 B200.

  python bench.py --gpus 1 --steps K --warmup W            this repo's CUDA engine through the C ABI
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   one rank per GPU
  python bench.py --impl reference ...                      the CPU restatement of the reference
                                                            (oracle/, "port"), all host threads,
                                                            on a bounded user sample of the same workload

A step = one whole RM2 job (jobs RM2-1..3 of M/rm/RM2Job.java:76-100) over the synthetic ML-20M-shaped
rating matrix: statistics, per-cluster matrices, scores and top-100 for all 138 493 users.
  value : users/s with the ratings already resident in HBM (timed: fy_rm2_run [+ NCCL gather])
  e2e   : users/s through the C ABI with HOST buffers: H2D of ratings + clustering, run, D2H of the
          packed top-N triples, every step
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rm2_users_scored_per_sec"
UNIT = "users/s"
LAMBDA = 0.1          # reference default, M/rmrecommender/RMRecommenderDriver.java:114
TOP_N = 100           # BASELINE.json configs: "RM2 top-100"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    ncu --set full capture (profiles/r01_ncu_traffic.json); None when no capture exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
            return float(json.load(f)[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hot = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(hot)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw))}


def workload_figures(r):
    """log-terms sum_u n_u * I_c(u) recomputed from the actual arrays."""
    n_u = np.bincount(r.user, minlength=r.n_users + 1)
    cl_of = np.zeros(r.n_users + 1, np.int64)
    cl_of[r.cl_user] = r.cl_cluster
    key = cl_of[r.user] * (int(r.item.max()) + 1) + r.item
    uniq = np.unique(key)
    i_c = np.bincount(uniq // (int(r.item.max()) + 1), minlength=r.n_clusters)
    terms = float(np.sum(n_u[r.cl_user].astype(np.float64) * i_c[r.cl_cluster]))
    return terms, i_c, n_u


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle ("port" of M/rm/AbstractRM2Reducer.java) on a bounded sample
# ------------------------------------------------------------------------------------------------
_CPU_CAL = {}


def cpu_baseline(r, budget_s=8.0, seed=12345):
    """The oracle ("port" of M/rm/AbstractRM2Reducer.java:129-233,321-371) on a bounded, seeded sample of users
    of one cluster, all host threads; users/s extrapolated by inner-loop work to the workload's mean user."""
    from oracle import rm2_oracle as orc
    cores = os.cpu_count() or 1
    terms_total, i_c, n_u = workload_figures(r)
    ksz = r.cluster_size.astype(np.float64)
    mean_inner = float(np.mean(ksz[r.cl_cluster] * n_u[r.cl_user] * i_c[r.cl_cluster]))
    # sample users of one (seeded) cluster so that one P cache is built, like one reduce() call
    rng = np.random.default_rng(seed)
    c = int(rng.integers(0, r.n_clusters))
    members = rng.permutation(r.cl_user[r.cl_cluster == c])
    inner = lambda us: float(np.sum(ksz[c] * n_u[us] * i_c[c]))
    t0 = time.time()
    if "rate" not in _CPU_CAL:
        # calibrate once on the lightest users (they run ~3x faster per unit of work than the average user)
        light = members[np.argsort(n_u[members])[:max(2, min(cores, len(members)))]]
        cal = orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, LAMBDA, r.n_items, TOP_N,
                      mode=orc.MODE_LITERAL_FAST, threads=cores, only_users=light)
        _CPU_CAL["rate"] = inner(light) / max(cal["seconds"], 1e-3) / 3.0   # inner iterations / s with all cores
    target = _CPU_CAL["rate"] * budget_s
    sample, acc = [], 0.0
    for u in members:
        sample.append(int(u)); acc += float(ksz[c] * n_u[u] * i_c[c])
        if acc >= target and len(sample) >= min(cores, 8):
            break
    cal_s = time.time() - t0
    out = orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, LAMBDA, r.n_items, TOP_N,
                  mode=orc.MODE_LITERAL_FAST, threads=cores, only_users=np.array(sample, np.int32))
    secs = max(out["seconds"], 1e-6)
    users_per_s_sample = len(sample) / secs
    # extrapolate by work: the sample's inner-iteration rate applied to the workload's mean user
    users_per_s = (inner(np.array(sample)) / secs) / mean_inner
    return {"value": users_per_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d users of cluster %d (K=%d, I_c=%d), scoring loops only (%.1f s wall, %d threads, "
                      "oracle MODE_LITERAL_FAST = the Java loop nest's arithmetic and order on a transposed P cache); "
                      "users/s extrapolated by inner-loop work K*n_u*I_c to the workload's mean user "
                      "(sample itself: %.3f users/s); Hadoop/JVM overheads not modelled"
                      % (len(sample), c, int(ksz[c]), int(i_c[c]), secs, cores, users_per_s_sample),
            "calibration_s": cal_s}


def run_reference_arm(args, r, workload):
    """--impl reference: the CPU restatement timed on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    budget = max(2.0, min(8.0, 60.0 / max(1, args.steps + args.warmup)))
    for s in range(args.warmup + args.steps):
        last = cpu_baseline(r, budget_s=budget, seed=12345 + s)
        if s >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    last = dict(last); last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r.n_users / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload, "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the reference (no JVM in this image); ms_per_step = extrapolated whole job"}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter (e.g. NCCL's version banner, printed with printf) must not share stdout with the
    ONE JSON line: point fd 1 at stderr and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(line, flush=True)
    else:
        os.write(_REAL_STDOUT, (line + "\n").encode())


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ml-20m")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--score-mode", type=int, default=0, help="0 = hi-word stream + exact re-score (default), 1 = fp64 stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from filmyou_core_b200 import datagen
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return
    t_gen = time.time()
    r = datagen.generate(args.workload)
    t_gen = time.time() - t_gen
    workload = {"workload": "RM2 top-%d, %s shape (synthetic, seed %d): %d users x %d items, %d ratings, %d clusters, lambda=%g"
                            % (TOP_N, args.workload, r.seed, r.n_users, r.n_items, r.nnz, r.n_clusters, LAMBDA),
                "sha256": r.sha256(), "l2": "inputs larger than L2 (per-cluster H is GBs; no flush needed)",
                "parallelism": "users sharded over %d GPU(s), ratings replicated" % world}
    if args.impl == "reference":
        return run_reference_arm(args, r, workload)

    import torch
    import torch.distributed as dist
    import filmyou_core_b200 as fy
    from filmyou_core_b200 import sharding

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = fy.Rm2Engine(lam=LAMBDA, number_of_items=r.n_items, top_n=TOP_N, device=local_rank,
                       shard_rank=rank, shard_count=world, score_mode=args.score_mode)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)          # torch.cuda.Event then times the launching stream

    def gather():
        if world == 1:
            return None
        d = eng.results_device()
        local = {k: torch.as_tensor(v, device=dev) for k, v in d.items()} if eng.result_count() else \
                {k: torch.empty(0, device=dev) for k in d}
        return sharding.gather_results(local, device=dev)

    # ---- device-resident: ratings already in HBM ----
    eng.set_ratings(r.user, r.item, r.score)
    eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
    for _ in range(args.warmup):
        eng.run(); gather()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    profs, launches = [], 0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        eng.run(); gather()
        p = eng.profile(); profs.append(p); launches += p["kernel_launches"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    scored = torch.tensor([float(eng.users_scored())], dtype=torch.float64, device=dev)
    sums = torch.tensor([sum(p["ms_score"] for p in profs), sum(p["score_bytes"] for p in profs),
                         sum(p["ms_gram"] for p in profs), sum(p["ms_index"] for p in profs),
                         sum(p["ms_topn"] for p in profs), float(launches),
                         float(sum(p["score_launches"] for p in profs)), sum(p["ms_refine"] for p in profs),
                         float(profs[-1]["bytes_per_term"]), float(sum(p["exact_rerun"] for p in profs)),
                         float(profs[-1]["score_kernel"])],
                        dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(scored, op=dist.ReduceOp.SUM)
        all_sums = [torch.zeros_like(sums) for _ in range(world)]
        dist.all_gather(all_sums, sums)
    else:
        all_sums = [sums]
    ms_per_step = float(ms.item()) / args.steps
    users = float(scored.item())
    value = users / (ms_per_step * 1e-3)

    # ---- end to end: host buffers in, host triples out, every step ----
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_user, h_item, h_score = pin(r.user), pin(r.item), pin(r.score)
        n_max = r.n_users * TOP_N
        outp = dict(user=torch.empty(n_max, dtype=torch.int32).pin_memory(), item=torch.empty(n_max, dtype=torch.int32).pin_memory(),
                    score64=torch.empty(n_max, dtype=torch.float64).pin_memory(), score32=torch.empty(n_max, dtype=torch.float32).pin_memory(),
                    cluster=torch.empty(n_max, dtype=torch.int32).pin_memory())
        outn = {k: v.numpy() for k, v in outp.items()}

        def e2e_step():
            eng.set_ratings(h_user.numpy(), h_item.numpy(), h_score.numpy())
            eng.set_clustering(r.cl_user, r.cl_cluster, r.cluster_size)
            eng.run()
            g = gather()
            if world == 1:
                res = eng.results(out=outn)
                return len(res["user"])
            n = g["user"].numel()               # every rank reads the gathered triples back to the host
            for k in outp:
                outp[k][:n].copy_(g[k], non_blocking=True)
            torch.cuda.synchronize()
            return n
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_out = 0
        for _ in range(args.steps):
            n_out = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_s = float(dt.item()) / args.steps
        e2e = {"value": users / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(r.nnz * 12 + r.n_users * 8 + r.n_clusters * 4),
               "d2h_bytes_per_step": int(n_out * 24), "ms_per_step": e2e_s * 1e3,
               "timing": "host wall clock around set_ratings+set_clustering+run+results, max over ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_score), CUDA events on the launching stream ----
    peak, peak_src = measured_peaks()
    per_rank = []
    for s in all_sums:
        s = s.tolist()
        per_rank.append({"ms_score": s[0], "score_bytes": s[1], "ms_gram": s[2], "ms_index": s[3], "ms_topn": s[4],
                         "launches": s[5], "score_launches": s[6], "ms_refine": s[7], "bytes_per_term": s[8],
                         "exact_rerun": s[9], "score_kernel": int(s[10])})
    worst = max(per_rank, key=lambda x: x["ms_score"])
    achieved = worst["score_bytes"] / (worst["ms_score"] * 1e-3) / 1e9 if worst["ms_score"] > 0 else 0.0
    hi = worst["bytes_per_term"] == 4.0
    kname = ("k_score", "k_score_hi", "k_score_f32")[worst["score_kernel"]]
    roofline = {"kernel": "fy::" + kname, "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(kname),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": worst["score_bytes"] / max(worst["score_launches"], 1),
                "avg_launch_ms": worst["ms_score"] / max(worst["score_launches"], 1),
                "definition": "%d B per (user, candidate, rated item) log-term = one %s element of H streamed; "
                              "sum over launches / sum of launch durations (CUDA events on the launching stream); "
                              "frac > 1 = rows re-used out of L2 (traffic = DRAM bytes per launch from the ncu capture "
                              "of one ML-20M-sized cluster, profiles/)" % (4 if hi else 8, "hi-word (4-byte)" if hi else "fp64"),
                "stage_ms_per_step": {k: worst[k] / args.steps for k in ("ms_index", "ms_gram", "ms_score", "ms_topn", "ms_refine")},
                "stages_overlap": "H build, score and top-N/refine run on three streams; stage times are per stream and overlap",
                "exact_reruns": worst["exact_rerun"]}
    # The plane rows are served out of L2 (traffic << algorithmic bytes), so the unit that actually bounds this kernel is
    # the L2 -> SM path: /opt/skills/guides/B300_MICROARCH.md measures a full-chip LTS throughput cap of ~6300 B/clk
    # (same L2 on B200); at the SM clock sampled during the timed region that is the ceiling reported here.
    if clocks and clocks.get("sm_mhz"):
        cap = 6300.0 * clocks["sm_mhz"] * 1e6 / 1e9
        roofline["l2"] = {"cap_bytes_per_clk": 6300, "sm_mhz": clocks["sm_mhz"], "cap": cap, "unit": "GB/s", "frac": achieved / cap,
                          "source": "B300_MICROARCH.md 'LTS throughput cap ~6300 B/cyc full-chip'; achieved = the same algorithmic bytes / launch time"}
    # second kernel of the step, for the record: k_build_H writes the fp64 plane (+ the 4-byte plane in auto mode)
    gram_bytes = sum(p["gram_bytes"] for p in profs) * (1.5 if hi else 1.0)
    gram_ms = sum(p["ms_gram"] for p in profs)
    roofline["secondary"] = {"kernel": "fy::k_build_H", "bound": "hbm", "unit": "GB/s",
                             "achieved": gram_bytes / (gram_ms * 1e-3) / 1e9 if gram_ms > 0 else 0.0, "peak": peak,
                             "frac": (gram_bytes / (gram_ms * 1e-3) / 1e9 / peak) if gram_ms > 0 else 0.0,
                             "definition": "bytes of H written (12 B per element with the 4-byte plane, 8 B without) / time of "
                                           "the H-build stream segments of rank 0 (CUDA events; the stage overlaps the score kernel)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(sum(p["launches"] for p in per_rank)), "roofline": roofline,
            "users_scored": users, "datagen_s": t_gen}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(r)
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
