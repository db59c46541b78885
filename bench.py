#!/usr/bin/env python3
"""bench.py -- queries/sec of the Infidex search path on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
  python bench.py --impl reference ...                      # the reference algorithm on the host cores (oracle port; the
                                                            # C# reference itself cannot run here -- no dotnet in the image)
A "step" is one pass of the hot path over one batch of synthetic queries (workload below). `value` times the
device-resident batch (ifx_batch_run, CUDA events inside the library); `e2e` times the reference-facing call
ifx_search_batch with host buffers (host->device query upload + device->host result download inside the region).
N > 1: one process per GPU (torchrun), each rank holds a replica of the index and runs its own batch per step
(weak scaling, queries are independent); per-rank results are all-gathered to rank 0 with NCCL. The doc-id-sharded
100M-doc configuration (configs[4]) is not built this round -- see DESIGN.md "multi-GPU".
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

WORKLOADS = {
    # BASELINE.json configs[1]: 1M synthetic single-field docs, 1k-query batch, top-10
    "c2": dict(n_docs=1_000_000, nq=1000, multi=False, vocab=400_000, label="configs[1]: 1M single-field docs, 1k-query batch, top-10, coverage depth 500"),
    # BASELINE.json configs[2]/[3]: 10M docs title(High)+description(Low), 10k-query batch (+ filter/facets with --filter)
    "c3": dict(n_docs=10_000_000, nq=10_000, multi=True, vocab=400_000, label="configs[2]: 10M multi-field docs, 10k-query batch, top-10"),
    "tiny": dict(n_docs=50_000, nq=200, multi=False, vocab=50_000, label="smoke workload (not a benchmark)"),
}


def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append((float(f[0]), float(f[1]), f[2], f[3], f[4], f[5]))
        except Exception:
            pass
        stop.wait(0.2)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_corpus(wl, seed_shift=0):
    from infidex_b200 import synth
    vocab = synth.make_vocab(wl["vocab"])
    docs = synth.gen_docs(wl["n_docs"], vocab, with_description=wl["multi"])
    schema, cols = synth.schema_and_columns(docs, wl["multi"])
    return vocab, docs, schema, cols


def run_reference(args, wl, rank, world):
    """Reference arm: the reference's algorithm (oracle port) on the host cores, all threads, bounded sample per step."""
    if rank != 0:
        return
    from infidex_b200 import synth
    from oracle.oracle import Field as OField
    from oracle.oracle import OracleEngine
    vocab, docs, schema, cols = make_corpus(wl)
    orc = OracleEngine([OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema])
    orc.index_columns(docs["keys"], cols)
    cores = os.cpu_count() or 1
    sample = min(wl["nq"], args.ref_sample)
    flt = None
    if args.filter:
        import infidex_b200 as ib
        flt = ib.Filter.Parse("year >= 2000 AND rating > 7.0").bytecode()
    times = []
    for s in range(args.warmup + args.steps):
        qs = synth.gen_queries(sample, docs, vocab, seed=synth.SEED + s)
        t0 = time.perf_counter(); orc.search_batch(qs, 10, 500, True, flt, threads=cores); dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    total = sum(times); qps = sample * len(times) / total
    line = {"impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "batch": sample, "filter": bool(args.filter)},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": "%d-query sample of the batch per step, oracle (C++ restatement of the C# reference), %d host threads" % (sample, cores)},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--filter", action="store_true", help="configs[3]: Filter.Parse('year >= 2000 AND rating > 7.0') + EnableFacets (multi-field workloads)")
    ap.add_argument("--ref-sample", type=int, default=1000, help="queries per step of the CPU arms (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)

    import infidex_b200 as ib
    from infidex_b200 import dist as ifxd
    from infidex_b200 import synth
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t_setup = time.time()
    eng = ib.SearchEngine.CreateDefault(device=local)
    if wl["n_docs"] > 2_000_000:          # large corpora are generated and indexed chunk by chunk
        vocab = synth.make_vocab(wl["vocab"])
        cc = synth.ChunkedCorpus(wl["n_docs"], vocab, wl["multi"], chunk=250_000)
        eng.IndexChunks(cc.schema, cc.chunks())     # numpy generation is fastest single-threaded (measured)
        docs = cc.docs_for_queries(); schema = cc.schema; cols = None; text_mb = cc.text_chars * 2 / 1e6
    else:
        vocab, docs, schema, cols = make_corpus(wl)
        eng.IndexColumns(docs["keys"], schema, cols); text_mb = docs["title"][1][-1] * 2 / 1e6
    t_index = time.time() - t_setup
    flt = ib.Filter.Parse("year >= 2000 AND rating > 7.0") if args.filter else None
    n_total = args.warmup + args.steps
    batches = []
    for s in range(n_total):   # a distinct batch per step and per rank
        qs = synth.gen_queries(wl["nq"], docs, vocab, seed=ifxd.rank_batch_seed(synth.SEED, s, rank))
        qq = []
        for t in qs:
            x = ib.Query(t, 10); x.Filter = flt; x.EnableFacets = bool(flt); qq.append(x)
        batches.append(qq)
    t_setup = time.time() - t_setup

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- value: device-resident batches, CUDA-event timing inside the library ------------------------------------------------
    handles = [eng.UploadBatch(b) for b in batches]
    stop = threading.Event(); clk = []
    th = threading.Thread(target=clocks_sampler, args=(stop, clk, local), daemon=True)
    agg = {k: 0.0 for k in ("ms_total", "ms_prepare", "ms_expand", "ms_stage1", "ms_wordmatch", "ms_stage2", "ms_final")}
    algo = 0; launches = 0; q_max = 0.0; q_sum = 0.0
    for s in range(n_total):
        if s == args.warmup:
            barrier(); th.start()
        eng.FlushL2()                       # cold L2 before every step (outside the event-timed region)
        st = eng.RunBatch(handles[s])
        if s >= args.warmup:
            for k in agg:
                agg[k] += getattr(st, k)
            algo += st.algo_bytes_stage1; launches += st.kernel_launches; q_max = max(q_max, st.s1_query_ms_max); q_sum += st.s1_query_ms_sum
    barrier()
    # ---- e2e: host buffers in / out through the C-ABI call ifx_search_batch (query upload + result download inside the region) --
    packed = [eng.PackBatch(b) for b in batches]
    e2e_t = 0.0; h2d = d2h = 0
    for s in range(n_total):
        eng.FlushL2()
        st = ib.Stats(); t0 = time.perf_counter(); eng.SearchPacked(packed[s], st); dt = time.perf_counter() - t0
        if s >= args.warmup:
            e2e_t += dt; h2d, d2h = st.h2d_bytes, st.d2h_bytes
    res_status = packed[-1]["bufs"]["status"]
    stop.set()
    bad = int(((res_status & ~8) != 0).sum())
    dev_ms = agg["ms_total"]; step_ms = dev_ms / args.steps
    if dist is not None:
        import torch
        dev_ms, e2e_t = ifxd.max_over_ranks(dist, [dev_ms, e2e_t], device="cuda"); step_ms = dev_ms / args.steps
        gathered = ifxd.gather_results(dist, packed[-1]["bufs"]["keys"], device="cuda")      # per-batch result exchange over NCCL
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    nq_all = wl["nq"] * world
    value = nq_all * args.steps / (dev_ms / 1e3)
    e2e = nq_all * args.steps / e2e_t
    peak, peak_src = measured_peak()
    s1_ms = agg["ms_stage1"] / args.steps; s1_bytes = algo / args.steps
    achieved = (s1_bytes / 1e9) / (s1_ms / 1e3) if s1_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "stage1_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload)
    clocks = None
    if clk:
        sm = sorted(c[0] for c in clk); reasons = set()
        for c in clk:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        clocks = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(c[1] for c in clk), "reasons": sorted(reasons), "samples": len(clk)}
    line = {"metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "batch_per_gpu": wl["nq"], "filter": bool(flt), "parallelism": "replica x%d (query-parallel)" % world,
                       "l2": "256 MiB L2 flush before every timed step; the index (text alone %.0f MB) also exceeds the 126 MB L2" % text_mb, "index_build_s": round(t_index, 1),
                       "setup_s": round(t_setup, 1), "bad_status": bad},
            "phases_ms_per_step": {k: round(v / args.steps, 3) for k, v in agg.items()},
            "roofline": {"bound": "hbm", "kernel": "k_stage1", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algo_bytes_per_launch": s1_bytes, "ms_per_launch": s1_ms, "peak_source": peak_src,
                         "longest_query_ms": q_max, "sum_query_ms_per_launch": q_sum / args.steps},
            "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks}
    if not args.no_cpu_baseline and world == 1 and cols is not None:
        line["cpu_baseline"] = cpu_baseline(wl, vocab, docs, schema, cols, flt, args)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(wl, vocab, docs, schema, cols, flt, args):
    """The oracle (CPU restatement of the reference) timed on this box's host cores over a bounded sample of the same workload."""
    from infidex_b200 import synth
    from oracle.oracle import Field as OField
    from oracle.oracle import OracleEngine
    orc = OracleEngine([OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema])
    orc.index_columns(docs["keys"], cols)
    cores = os.cpu_count() or 1
    sample = min(wl["nq"], args.ref_sample)
    qs = synth.gen_queries(sample, docs, vocab, seed=synth.SEED + 1)
    fb = flt.bytecode() if flt else None
    orc.search_batch(qs[: max(8, sample // 20)], 10, 500, True, fb, threads=cores)       # warm-up
    t0 = time.perf_counter(); orc.search_batch(qs, 10, 500, True, fb, threads=cores); dt = time.perf_counter() - t0
    t1 = time.perf_counter(); orc.search_batch(qs[: max(50, sample // 10)], 10, 500, True, fb, threads=1); dt1 = time.perf_counter() - t1
    return {"value": sample / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "single_thread_value": max(50, sample // 10) / dt1,
            "sample": "%d queries of the step-1 batch, oracle (C++ restatement of the C# reference; dotnet absent), %d threads" % (sample, cores)}


if __name__ == "__main__":
    main()
