#!/usr/bin/env python
"""bench.py -- read pairs/s aligned + piled up (+ called + classified) on B200, BASELINE.json's metric.

A STEP is one pass of the whole read-level hot path over one sample's batch of synthetic read pairs:
seeding/chaining -> ksw_extend2 rounds -> pairing + CIGAR -> pileup counts -> (N>1: NCCL all-reduce of the
int32 count tensor) -> SNP calls -> TP/FP/FN match against the strain-difference truth set.

Workload (config.workload): BASELINE.json configs[1], the 10-sample TB40E:AD169 abundance-ratio series,
2,000,000 synthetic 2x150 bp pairs per sample per GPU; step i runs sample i % 10.  With N GPUs each rank
takes pairs [r*P, (r+1)*P) of an N*P-pair sample (weak scaling), counts are merged with one all-reduce.

  value : whole-job pairs/s with the reads already resident in HBM (timed with CUDA events, max over ranks)
  e2e   : the same through the host-buffer C-ABI call (qm_sample_add_pairs_host): pinned host reads -> H2D ->
          pipeline -> calls D2H, all inside the timed region
  roofline / stages : per-stage CUDA-event times measured live by the library's stage timers (same run)
  cpu_baseline : the oracle port (oracle/, CPU restatement of bwa-mem extension + bcftools counting, OpenMP on
          all host cores) on a bounded sample of the same workload -- a reported baseline, not the target
  --impl reference : that CPU path as its own arm (the upstream binaries are not in the image: kind "port")
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "read_pairs_per_s_aligned_piledup"
UNIT = "pairs/s"
N_SAMPLES = 10
ALGO_BYTES_PER_PAIR_PILEUP = 424          # 2 x (38 + 150 + 16 + 8)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=2_000_000, help="pairs per sample per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def config_dict(args, n_gpus):
    return {"workload": "cfg2: 10-sample TA-* abundance-ratio series (TB40E:AD169 1:0 ... 0:1), synthetic 2x150 bp pairs, "
                        "step i = sample i%10, reference AD169 (TB40E for TA-1-0)",
            "pairs_per_step_per_gpu": args.pairs, "read_len": 150, "global_pairs_per_step": args.pairs * n_gpus,
            "parallelism": f"reads sharded over {n_gpus} GPU(s), int32 count tensor all-reduced" if n_gpus > 1 else "1 GPU",
            "l2_policy": "inputs larger than L2: 1.2 GB of reads per step vs 126 MB L2; every step runs another sample"}


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def cpu_arm(args, steps, warmup, sample_pairs):
    """the oracle port on host cores; each step = `sample_pairs` pairs of sample i % 10.  -> (pairs/s, ms/step)"""
    import numpy as np
    from quasimodo_b200 import workloads
    from oracle import qmo_py
    refs, times = {}, []
    for i in range(warmup + steps):
        W = workloads.config2(i % N_SAMPLES, sample_pairs)
        key = W.ref_stems[0]
        if key not in refs:
            refs[key] = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
        codes, quals, _, _ = W.simulate_host(0, sample_pairs)
        lens = np.full(2 * sample_pairs, W.params.read_len, np.int32)
        t0 = time.perf_counter()
        qmo_py.run_sample(refs[key], codes, quals, lens)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return sample_pairs * len(times) / tot, tot / len(times) * 1e3


def calibrated_cpu_sample(args):
    """pick a sample size that costs about args.cpu_seconds of CPU wall time"""
    rate, _ = cpu_arm(args, 1, 0, 20_000)
    n = int(max(20_000, min(args.pairs, rate * args.cpu_seconds)))
    return n - n % 1000


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import qmo_py
    qmo_py.build()
    cores = qmo_py.n_threads()
    per_step = max(20_000, int(calibrated_cpu_sample(args) / max(1, args.steps + args.warmup) * 4))
    per_step -= per_step % 1000
    value, ms = cpu_arm(args, args.steps, args.warmup, per_step)
    sample = f"{per_step} pairs per step (prefix of each step's sample), oracle port: seeding+extension+mate rescue+pairing+CIGAR+pileup"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "int32", "data": "synthetic", "config": config_dict(args, args.gpus),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "bwa/samtools/bcftools are not in the image and not vendored: the CPU arm is the repo's C restatement "
                   "(oracle/, OpenMP over reads), not the upstream binaries"}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from quasimodo_b200 import Context, _lib, evaluate, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    P, K, Wm = args.pairs, args.steps, args.warmup
    ctx = Context(local)                              # raises without the CUDA library / a B200: no CPU fallback
    st = torch.cuda.current_stream().cuda_stream
    n_used = min(N_SAMPLES, K + Wm)

    # ---- per-sample state: workload, index + qm_sample per reference, truth keys, resident reads ----
    wl = [workloads.config2(i, P * world) for i in range(n_used)]
    idx, smp, tkeys_dev, tkeys = {}, {}, {}, {}
    for W in wl:
        key = W.ref_stems[0]
        if key not in idx:
            idx[key] = ctx.index(W.ref, 31)
            smp[key] = ctx.sample(idx[key])
    truth_path = os.path.join(ROOT, "quasimodo_b200", "data", "truth", "TA.maskrepeat.variants.vcf.gz")
    import gzip
    with gzip.open(truth_path, "rt") as fh:
        tk = []
        for ln in fh:
            f = ln.rstrip("\n").split("\t")
            if len(f) >= 5 and f[3] in "ACGT" and f[4] in "ACGT" and len(f[3]) == 1 and len(f[4]) == 1 and f[1].isdigit():
                tk.append(int(evaluate.snp_key(f[1], f[3], f[4])))
    tkeys = np.array(tk, dtype=np.uint64)
    d_tkeys = torch.from_numpy(tkeys.view(np.int64)).to(dev)
    L = 150
    d_lens = torch.full((2 * P,), L, dtype=torch.int32, device=dev)
    d_lens_pre = torch.full((2 * _lib.PESTAT_PAIRS,), L, dtype=torch.int32, device=dev)
    d_reads = []
    d_prefix = []
    npre = min(_lib.PESTAT_PAIRS, P * world)
    for W in wl:
        g = torch.from_numpy(W.src_codes).to(dev)
        c = torch.empty((2 * P, L), dtype=torch.uint8, device=dev)
        q = torch.empty((2 * P, L), dtype=torch.uint8, device=dev)
        ctx.simulate_pairs(W, rank * P, P, g, c, q, st)
        d_reads.append((c, q))
        if rank != 0:                                  # the sample's designated insert-size prefix (pairs 0..65535)
            pc = torch.empty((2 * npre, L), dtype=torch.uint8, device=dev)
            pq = torch.empty((2 * npre, L), dtype=torch.uint8, device=dev)
            ctx.simulate_pairs(W, 0, npre, g, pc, pq, st)
            d_prefix.append(pc)
            del pq
        else:
            d_prefix.append(None)
    torch.cuda.synchronize()
    max_calls = 1 << 18
    d_calls = torch.empty(max_calls * 40, dtype=torch.uint8, device=dev)
    d_ckeys = torch.empty(max_calls, dtype=torch.int64, device=dev)
    d_cflags = torch.zeros(max_calls, dtype=torch.uint8, device=dev)
    d_tflags = torch.zeros(len(tkeys), dtype=torch.uint8, device=dev)
    copt = _lib.default_call_opt()
    import ctypes as C
    lib = _lib.lib()
    results = {}

    def finish_sample(i, s, key):
        """all-reduce (N>1), call SNPs, classify against the truth set; returns (n_calls, tp, fp, fn)"""
        if world > 1:
            dist.all_reduce(s.counts_tensor(), op=dist.ReduceOp.SUM)
        if rank != 0:
            return None
        n = C.c_int64()
        rc = lib.qm_call_snps(ctx._h, idx[key]._h, C.byref(copt), C.c_void_p(s.counts_ptr()), C.c_void_p(d_calls.data_ptr()),
                              max_calls, C.byref(n), C.c_void_p(st))
        if rc:
            raise RuntimeError(lib.qm_last_error(ctx._h).decode())
        nc = n.value
        calls = d_calls[:nc * 40].view(torch.int32).view(nc, 10)
        # key = (pos+1) << 8 | ref << 4 | alt ; ref/alt are bytes 0/1 of the third int32
        ra = calls[:, 2].to(torch.int64)
        d_ckeys[:nc] = ((calls[:, 1].to(torch.int64) + 1) << 8) | ((ra & 0xff) << 4) | ((ra >> 8) & 0xff)
        pure = wl[i].name.endswith(("-1-0", "-0-1"))
        if pure or nc == 0:
            return nc, 0, nc, 0
        rc = lib.qm_eval_match(ctx._h, C.c_void_p(d_ckeys.data_ptr()), nc, C.c_void_p(d_tkeys.data_ptr()), len(tkeys),
                               C.c_void_p(d_cflags.data_ptr()), C.c_void_p(d_tflags.data_ptr()), C.c_void_p(st))
        if rc:
            raise RuntimeError(lib.qm_last_error(ctx._h).decode())
        tp = int(d_cflags[:nc].sum().item())
        fn = int(len(tkeys) - d_tflags.sum().item())
        return nc, tp, nc - tp, fn

    def step_resident(i):
        W = wl[i % n_used]
        key = W.ref_stems[0]
        s = smp[key]
        s.reset(st)
        c, q = d_reads[i % n_used]
        if rank != 0:
            s.estimate_pestat(d_prefix[i % n_used], d_lens_pre[:2 * npre], st)
        s.add_pairs(c, q, d_lens, pair_id0=rank * P, stream=st)
        r = finish_sample(i % n_used, s, key)
        if r is not None:
            results[W.name] = r
        return s.stats(st)[1]                          # executed ksw_extend2 cells of this step (8-byte read-back)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up, then EXACTLY K timed steps between barriers ----
    for i in range(Wm):
        step_resident(i)
    barrier()
    ctx.profile_collect()
    ctx.profile_enable(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cells_total = 0
    barrier()
    e0.record()
    for i in range(K):
        cells_total += step_resident(Wm + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    stage_ms, stage_launch = ctx.profile_collect()
    ctx.profile_enable(False)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = P * world * K / ms * 1e3

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        h_codes = torch.empty((2 * P, L), dtype=torch.uint8).pin_memory()
        h_quals = torch.empty((2 * P, L), dtype=torch.uint8).pin_memory()
        h_lens = torch.full((2 * P,), L, dtype=torch.int32).pin_memory()
        h_calls = torch.empty(max_calls * 40, dtype=torch.uint8).pin_memory()
        e2e_times, d2h = [], 0

        def step_host(i):
            W = wl[i % n_used]
            key = W.ref_stems[0]
            s = smp[key]
            s.reset(st)
            if rank != 0:
                s.estimate_pestat(d_prefix[i % n_used], d_lens_pre[:2 * npre], st)
            torch.cuda.synchronize()
            s.add_pairs_host(h_codes, h_quals, h_lens, pair_id0=rank * P)
            r = finish_sample(i % n_used, s, key)
            nb = 0
            if r is not None:
                nb = r[0] * 40
                h_calls[:nb].copy_(d_calls[:nb], non_blocking=True)
            torch.cuda.synchronize()
            return nb + 8

        for i in range(Wm + K):
            c, q = d_reads[i % n_used]
            h_codes.copy_(c)
            h_quals.copy_(q)
            barrier()
            t0 = time.perf_counter()
            nb = step_host(i)
            barrier()
            dt = time.perf_counter() - t0
            if i >= Wm:
                e2e_times.append(dt)
                d2h = max(d2h, nb)
        te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": P * world * K / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(2 * (2 * P * L) + 4 * 2 * P), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": float(te.item()) / K * 1e3,
               "api": "qm_sample_add_pairs_host (+ qm_call_snps, qm_eval_match), pinned host buffers"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + every stage ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    dpx_gops, _ = ctx.dpx_peak(1, 4096)                 # measured now: G lane-instr/s of __viaddmax_s16x2_relu
    gcups_peak = dpx_gops * 2.0 / 9.0                   # 2 packed cells per lane-instr, 9 DPX-class instr per cell
    traffic = None
    try:                                                   # per-launch DRAM bytes of the extension kernel from the committed ncu capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ext2_traffic.json")))["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    ext_ms = stage_ms["extend"]
    dom = max(("seed_chain", "advance", "extend", "pair_cigar", "pileup"), key=lambda k: stage_ms[k])
    gcups = cells_total / ext_ms / 1e6 if ext_ms > 0 else 0.0
    roofline = {"kernel": "ext2_kernel<CAP> + ext_kernel<C> (batched ksw_extend2)", "bound": "int-issue (DPX), not hbm/tensor",
                "achieved": gcups, "peak": gcups_peak, "unit": "GCUPS", "frac": gcups / gcups_peak if gcups_peak else None,
                "traffic": traffic, "traffic_how": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per ext2_kernel launch "
                                               "(profiles/ext2_traffic.json); the kernel is issue-bound, not DRAM-bound",
                "avg_launch_ms": ext_ms / max(1, stage_launch["extend"]),
                "work": f"{cells_total} executed ksw_extend2 cells in {K} steps ({cells_total / (P * K):.0f} cells/pair)",
                "peak_how": f"measured in this run: {dpx_gops:.0f} G lane-instr/s of viaddmax_s16x2_relu x 2 cells / 9 instr",
                "dominant_stage": dom}
    pile_ms = stage_ms["pileup"]
    pile_gbs = ALGO_BYTES_PER_PAIR_PILEUP * P * K / pile_ms / 1e6 if pile_ms > 0 else 0.0
    roofline_pileup = {"kernel": "pileup_kernel", "bound": "hbm", "achieved": pile_gbs, "peak": hbm_peak, "unit": "GB/s",
                       "frac": pile_gbs / hbm_peak, "traffic": None, "peak_how": hbm_src,
                       "work": f"{ALGO_BYTES_PER_PAIR_PILEUP} B/pair algorithmic x {P * K} pairs"}
    launches = int(sum(stage_launch.values()))
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
           "data": "synthetic", "config": config_dict(args, world), "clocks": clk, "e2e": e2e, "gpu_launches": launches,
           "roofline": roofline, "roofline_pileup": roofline_pileup,
           "stages_ms_per_step": {k: v / K for k, v in stage_ms.items()},
           "stage_launches": stage_launch,
           "results": {k: {"calls": v[0], "TP": v[1], "FP": v[2], "FN": v[3]} for k, v in results.items()}}

    if not args.no_cpu_baseline and world == 1:
        from oracle import qmo_py
        qmo_py.build()
        n = calibrated_cpu_sample(args)
        v, cms = cpu_arm(args, 1, 0, n)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": qmo_py.n_threads(), "kind": "port",
                               "sample": f"first {n} pairs of sample TA-1-0 ({cms / 1e3:.1f} s): oracle port of bwa-mem "
                                         "extension + mate rescue + pairing + CIGAR + bcftools-style counting, OpenMP over reads"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
