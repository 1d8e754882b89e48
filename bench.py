#!/usr/bin/env python
"""Benchmark of the baseband -> CADU hot path (BASELINE.json metric: baseband MS/s end to end, % of the HBM roofline).

  python bench.py --gpus N --steps K --warmup W                 headline: METOP AHRPT (BASELINE configs[2]), one stream per GPU
  python bench.py --config {c2,c3,c4,c5} ...                    the other BASELINE configurations (c3 = default)
  python bench.py --impl reference ...                          the reference's own CPU code (oracle/_ref) on the host cores

A step = one pass of the whole hot path over one batch of `2**log2_samples` synthetic samples per GPU.
  value : samples/s with the batch already resident in HBM when the timed region starts (device timed, max over ranks)
  e2e   : the same through the public C ABI with HOST (pinned) buffers: H2D of the batch and D2H of the result inside the timed region
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

# BASELINE.json configs -> synthetic signal (satdump_b200/synth.py), stream-seed index, what a step produces
WORKLOADS = {
    "c3": dict(sig="metop_ahrpt", idx=3, kind="chain", metric="baseband MS/s end-to-end IQ->CADU (METOP AHRPT)",
               label="METOP AHRPT QPSK + Viterbi r=3/4 + RS(255,223) I=4, cs16 @6 MS/s signal (BASELINE configs[2]), one stream per GPU"),
    "c2": dict(sig="bpsk_half", idx=2, kind="chain", metric="baseband MS/s end-to-end IQ->CADU (BPSK r=1/2 + RS)",
               label="BPSK + Viterbi r=1/2 + RS(255,223) I=4 (psk_demod + ccsds_conv_concat_decoder), cf32 @3 MS/s signal, 1.2 Msym/s (BASELINE configs[1]), one stream per GPU"),
    "c4": dict(sig="jpss_hrd", idx=4, kind="chain", metric="baseband MS/s end-to-end IQ->CADU (JPSS HRD OQPSK)",
               label="JPSS-HRD-type OQPSK + Viterbi r=1/2 + NRZ-M + RS(255,223) I=5, cs16 @30 MS/s signal, 15 Msym/s (BASELINE configs[3]), one stream per GPU"),
    "c5": dict(sig="dvbs2_front", idx=5, kind="demod", metric="baseband MS/s IQ->symbols (DVB-S2 front half AGC->RRC->M&M)",
               label="DVB-S2 front half AGC -> RRC -> M&M (dvbs2_demod), cs8, 45 Msym/s @90 MS/s (BASELINE configs[4]), one stream per GPU"),
}
FMT_BYTES = {"cf32": 8, "cs16": 4, "cs8": 2}


def measured_traffic(kernel, log2n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture (these kernels' traffic is
    linear in the batch, so other batch sizes are scaled from the captured one)."""
    for name in ("ncu_r2_traffic.json", "ncu_r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            if kernel in t["dram_bytes_per_launch"]:
                return int(t["dram_bytes_per_launch"][kernel] * 2.0 ** (log2n - t["batch_log2_samples"]))
        except Exception:
            pass
    return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed regions. In-process NVML (nvidia_ml_py) every 25 ms; spawning nvidia-smi
    (the fallback) costs ~0.5 s of driver-lock time per call and visibly slows a 200 ms timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.max_sm, self.mask, self.stop_flag, self.how = index, [], None, 0, False, "nvml"

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.sm.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mask |= int(reasons(h))
                time.sleep(0.025)
            return
        except Exception:
            self.how = "nvidia-smi"
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")]
                if r and r[0].isdigit():
                    self.sm.append(int(r[0]))
                    self.max_sm = int(r[1]) if r[1].isdigit() else self.max_sm
                    for i, bit in enumerate((0x8, 0x40, 0x20, 0x4)):
                        if len(r) > 2 + i and r[2 + i].lower().startswith("active"):
                            self.mask |= bit
            except Exception:
                pass
            time.sleep(0.5)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": [n for b, n in self.REASONS.items() if self.mask & b],
                "samples": len(sm), "source": self.how}


def make_workload(w, log2n, rank, device, esn0=None):
    import dataclasses
    from satdump_b200 import shard, synth
    cfg = synth.CONFIGS[w["sig"]]
    if esn0 is not None:
        cfg = dataclasses.replace(cfg, esn0_db=esn0)
    raw, clear = synth.make_signal(cfg, 1 << log2n, seed=shard.stream_seed(w["idx"], rank), device=device)  # stream = rank
    return cfg, raw, clear


def nsamples_of(raw, cfg):
    return raw.numel() if cfg.fmt == "cf32" else raw.numel() // 2


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def oracle_cfgs(O, cfg, w):
    extra = dict(clock_alpha=cfg.clock_alpha) if cfg.clock_alpha else {}
    dcfg = O.demod_cfg(cfg.samplerate, cfg.symbolrate, cfg.constellation if cfg.decoder != "none" else "none", cfg.rrc_alpha, cfg.pll_bw, cfg.fmt, **extra)
    if w["kind"] == "demod":
        return dcfg, None
    if cfg.decoder == "metop":
        return dcfg, O.metop_cfg(cfg.ber_thresold, cfg.outsync_after)
    return dcfg, O.ccsds_cfg(cfg.constellation, cfg.cadu_bytes * 8, cfg.ber_thresold, cfg.outsync_after, cfg.interleave, nrzm=cfg.nrzm,
                             rs_usecheck=cfg.rs_usecheck)


def cpu_pipelines(O, dcfg, fcfg, hraw, nstreams):
    """`nstreams` concurrent runs of the reference's threaded pipeline (each its own objects and ~8 threads, like N satdump processes)
    on the same input. Returns (wall seconds of the slowest, threads per pipeline, CADU/symbol bytes of stream 0)."""
    res = [None] * nstreams

    def one(i):
        res[i] = O.pipeline_timed(dcfg, fcfg, hraw)

    t0 = time.perf_counter()
    th = [threading.Thread(target=one, args=(i,)) for i in range(nstreams)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    return (res[0][0] if nstreams == 1 else wall), res[0][2], res[0][1]


def cpu_stage_rates(O, dcfg, fcfg, hraw, n):
    """Single-thread MS/s of every block of the reference (the `satdump dsp_bench` convention, dsp/benchmark/helpers.h:13-62), each
    fed the previous block's output, in baseband-sample equivalents."""
    out = {}
    m = min(n, 1 << 22)
    per = 1 if dcfg.format == 0 else 2
    d = O.Demod(dcfg)
    t = time.perf_counter()
    o = d.run(hraw[:m * per])
    out["demod_chain_sequential"] = m / (time.perf_counter() - t) / 1e6
    x = o["agc"]  # (stage inputs: the stage before's output)
    conv = (hraw[:m * per].astype(np.float32) / np.float32(32767 if dcfg.format == 1 else 127)).view(np.complex64) if dcfg.format else hraw[:m]
    for st, src in (("agc", conv), ("fir", o["agc"]), ("costas", o["fir"]), ("mm", o["fir"] if o["costas"] is None else o["costas"])):
        if st == "costas" and o["costas"] is None:
            continue
        t = time.perf_counter()
        O.run_stage(dcfg, st, src)
        out[st] = src.size / (time.perf_counter() - t) / 1e6
    if fcfg is not None:
        f = O.Fec(fcfg)
        t = time.perf_counter()
        f.run(o["soft"])
        out["viterbi_deframe_rs"] = m / (time.perf_counter() - t) / 1e6
    return {k: round(v, 2) for k, v in out.items()}


def cpu_baseline(w, cfg, hraw, n, repeats=3, detail=True):
    """The reference's own code on the host cores: generic (-O2, scalar VOLK shim) and native (-O3 + FMA) builds, median of `repeats`,
    one stream and floor(cores / 8) concurrent streams."""
    from oracle import port, ref, ref_native
    cores = os.cpu_count() or 1
    if not ref.available():
        dcfg, fcfg = oracle_cfgs(port, cfg, w)
        t0 = time.time()
        port.pipeline_run(dcfg, fcfg, hraw)
        v = n / (time.time() - t0) / 1e6
        return {"value": v, "unit": "MS/s", "cores": 1, "kind": "port", "sample": f"{n} samples, single-thread C restatement (oracle/_ref absent)"}, None
    res, cadu0 = {}, None
    for tag, O in (("generic_O2", ref), ("native_O3_fma", ref_native)):
        if not O.available():
            continue
        dcfg, fcfg = oracle_cfgs(O, cfg, w)
        one = []
        for _ in range(repeats):
            secs, threads, cadu = cpu_pipelines(O, dcfg, fcfg, hraw, 1)
            one.append(n / secs / 1e6)
        if cadu0 is None:
            cadu0 = cadu
        r = {"one_stream_MSps_median": round(sorted(one)[len(one) // 2], 2), "one_stream_MSps_runs": [round(x, 2) for x in one], "threads_per_stream": threads}
        k = max(1, cores // 8)
        if k > 1:
            many = []
            for _ in range(repeats):
                secs, _, _ = cpu_pipelines(O, dcfg, fcfg, hraw, k)
                many.append(k * n / secs / 1e6)
            r["concurrent_streams"] = k
            r["all_streams_MSps_median"] = round(sorted(many)[len(many) // 2], 2)
        if detail:
            r["stage_single_thread_MSps"] = cpu_stage_rates(O, dcfg, fcfg, hraw, n)
        res[tag] = r
    best_tag = max(res, key=lambda t: res[t].get("all_streams_MSps_median", res[t]["one_stream_MSps_median"]))
    best = res[best_tag]
    value = best.get("all_streams_MSps_median", best["one_stream_MSps_median"])
    k = best.get("concurrent_streams", 1)
    try:
        model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        model = "unknown"
    return ({"value": value, "unit": "MS/s", "cores": k * best["threads_per_stream"], "kind": "reference", "streams": k,
             "sample": f"{n} samples of the bench signal per stream; {k} concurrent stream(s) x {best['threads_per_stream']} threads (reference threading model, "
                       f"one thread per DSP block), build {best_tag}, generic non-SIMD VOLK shim (no VOLK library on this host); host: {cores} cores, {model}; "
                       f"median of {repeats}", "variants": res}, cadu0)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    log2n = min(args.log2_samples, 25)  # bounded sample of the same workload
    cfg, raw, _ = make_workload(w, log2n, 0, dev, args.esn0)
    hraw = raw.cpu().numpy()
    n = nsamples_of(raw, cfg)
    t0 = time.perf_counter()
    for _ in range(args.warmup):  # (page-in of the libraries and the input; the timed runs below each make `repeats` passes)
        pass
    cb, out0 = cpu_baseline(w, cfg, hraw, n, repeats=max(3, args.steps), detail=True)
    tot = time.perf_counter() - t0
    val = cb["value"]
    line = {"impl": "reference", "metric": w["metric"], "value": val, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": n * cb.get("streams", 1) / val / 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8",
            "data": "synthetic",
            "config": {"workload": w["label"], "samples_per_step_per_stream": n, "result_bytes_per_stream": int(out0.size) if out0 is not None else None,
                       "note": "bounded sample of the b200 arm's workload (same signal generator, stream of rank 0); value = all concurrent CPU streams",
                       "wall_seconds_total": round(tot, 1)},
            "cpu_baseline": cb, "e2e": {"value": val, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly one JSON line: whatever libraries print at C level (NCCL's version banner, ...) goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


# ------------------------------------------------------------------------------------------------ the two pipeline shapes
class ChainPipe:
    """psk_demod + decoder module through b200_chain_* (the int8 soft stream stays in HBM); a step's result = its CADUs."""

    def __init__(self, capi, cfg, n, local):
        self.max_soft = int(n * 2.0 / max(1.0, cfg.samplerate / cfg.symbolrate) * 1.02) + (1 << 20)
        self.ch = capi.Chain(capi.demod_cfg_for(cfg, n, device=local), capi.fec_cfg_for(cfg, self.max_soft, device=local))
        self.unit = cfg.cadu_bytes
        self.out_cap = 2 * (self.max_soft // 8) + (1 << 20)

    def reset(self):
        self.ch.reset()

    def push_device(self, ptr, n):
        self.ch.push_device(ptr, n)

    def push_host(self, ptr, n):
        self.ch.push_ptr(ptr, n)

    def prefetch(self, ptr, n):
        self.ch.prefetch_ptr(ptr, n)

    def pull_into(self, ptr, cap):
        return self.ch.pull_into(ptr, cap)

    def touch_device_result(self):
        self.ch.frames_device()

    def drain(self, ptr, cap):
        self.ch.sync()
        total = 0
        while True:
            nb = self.ch.pull_into(ptr, cap)
            if nb == 0:
                return total
            total += nb

    def set_pipelined(self, on):
        self.ch.set_pipelined(on)

    def span_begin(self):
        self.ch.span_begin()

    def span_end(self):
        return self.ch.span_end()

    def timing(self):
        return self.ch.timing()

    def launches(self):
        return sum(s["kernel_launches"] for s in self.ch.stats())

    def stats(self):
        d, f = self.ch.stats()
        return {"demod": {k: v for k, v in d.items() if k in ("costas_unconverged", "mm_unconverged", "repairs", "agc_clamped")},
                "fec": {k: v for k, v in f.items() if k in ("replays", "rs_failed", "rs_corrected", "viterbi_state", "deframer_state", "frames_out", "start_redone", "tb_serial",
                                                         "spec_steps", "tb_overlap")}}


class DemodPipe:
    """The demodulator alone through b200_demod_* (C5: AGC -> RRC -> M&M, no Costas loop, no decoder); a step's result = cf32 symbols."""

    def __init__(self, capi, cfg, n, local):
        self.d = capi.Demod(capi.demod_cfg_for(cfg, n, device=local))
        self.unit = 8
        self.out_cap = int(n / (cfg.samplerate / cfg.symbolrate) * 1.02 + 4096) * 8
        self.ms = 0.0

    def reset(self):
        self.d.reset()

    def push_device(self, ptr, n):
        self.d.push_device(ptr, n)
        self.ms += self.d.timing()["stages_sum"]

    def push_host(self, ptr, n):
        self.d.push_ptr(ptr, n)

    def prefetch(self, ptr, n):
        self.d.prefetch_ptr(ptr, n)

    def pull_into(self, ptr, cap):
        return self.d.pull_symbols_into(ptr, cap // 8) * 8

    def touch_device_result(self):
        pass

    def drain(self, ptr, cap):
        return 0

    def set_pipelined(self, on):
        pass

    def span_begin(self):
        self.ms = 0.0

    def span_end(self):
        return self.ms  # sum of the pushes' CUDA-event times (every push is synchronous: nothing overlaps between them)

    def timing(self):
        t = self.d.timing()
        t["push_events"] = t["stages_sum"]
        return t

    def launches(self):
        return self.d.stats()["kernel_launches"]

    def stats(self):
        return {"demod": {k: v for k, v in self.d.stats().items() if k in ("costas_unconverged", "mm_unconverged", "repairs", "agc_clamped", "symbols_out")}}


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--log2-samples", dest="log2_samples", type=int, default=29)
    ap.add_argument("--esn0", type=float, default=None, help="Es/N0 of the synthetic signal in dB (default: the configuration's, 10 dB for c3; 5.5 = RS doing work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.config]
    if args.impl == "reference":
        return run_reference(args, w)

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
    import torch
    import torch.distributed as dist
    from satdump_b200 import capi, shard
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    pin_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    cfg, raw, clear = make_workload(w, args.log2_samples, rank, dev, args.esn0)
    n = nsamples_of(raw, cfg)
    bps = FMT_BYTES[cfg.fmt]
    host = torch.empty(raw.shape, dtype=raw.dtype, pin_memory=True)
    host.copy_(raw)
    torch.cuda.synchronize()
    pipe = (ChainPipe if w["kind"] == "chain" else DemodPipe)(capi, cfg, n, local)
    unit = pipe.unit

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        pipe.reset()
        pipe.push_device(raw.data_ptr(), n)
        pipe.touch_device_result()
        return pipe.timing()["push_events"]  # ms, CUDA events on the pipeline's own streams (torch events cannot see them)

    out_host = torch.empty(pipe.out_cap, dtype=torch.uint8, pin_memory=True)  # the caller-owned result buffer

    def step_host():
        pipe.reset()
        pipe.push_host(host.data_ptr(), n)
        nb = pipe.pull_into(out_host.data_ptr(), out_host.numel())  # D2H of this step's result
        return out_host[:nb].numpy().copy()

    # warm-up (also the full-size correctness gate of the chain configurations: every CADU must be one of the transmitted frames, in order)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    res0 = step_host()
    nres = res0.size // unit
    frames_ok, first, parity_as_received = True, None, None
    if w["kind"] == "chain":
        fr = res0.reshape(-1, unit)
        # every CADU must be one of the transmitted frames, in order and without a gap. The comparison is over the ASM + the RS message
        # bytes: the decoder writes back only the corrected message (reedsolomon.cpp:96-104), a channel error in a codeword's parity bytes
        # stays in the CADU as received, in the reference as here (`cadus_with_parity_bytes_as_received` counts those frames). (At the
        # stress SNR frames RS could not repair differ from the transmitted ones, and rs_usecheck drops them: then nothing is compared.)
        msg = 4 + 223 * cfg.interleave if cfg.interleave else unit
        first = next((i for i in range(min(256, clear.shape[0])) if np.array_equal(clear[i, :msg], fr[0, :msg])), None) if nres else None
        strict = pipe.stats()["fec"]["rs_failed"] == 0
        frames_ok = nres > 0 and (not strict or (first is not None and first + nres <= clear.shape[0] and np.array_equal(fr[:, :msg], clear[first:first + nres, :msg])))
        parity_as_received = int((fr[:, msg:] != clear[first:first + nres, msg:]).any(axis=1).sum()) if (frames_ok and strict and first is not None) else None

    # PCIe ceiling of the e2e number: the same pinned batch copied host -> device alone (torch copy engine, CUDA events)
    dst = torch.empty_like(raw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        dst.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    h2d_gbps = 3 * host.numel() * host.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del dst

    sampler = ClockSampler(local)
    sampler.start()
    # ---- per-kernel / per-stage times: one synchronous push (each kernel alone on the GPU), CUDA events on the pipeline's own streams
    barrier()
    sync_ms = 0.0
    for _ in range(args.steps):
        sync_ms += step_dev()
    tim = pipe.timing()
    sync_ms /= args.steps
    # ---- value: inputs resident in HBM. Chain: pipelined (decoder one batch behind the demodulator on its own stream / thread, the
    # reference's one-thread-per-module model). Every step is a fresh stream (reset), K steps + drain inside the stopwatch.
    pipe.set_pipelined(True)
    for _ in range(2):  # warm the pipelined path (staging buffers, worker thread)
        pipe.reset()
        pipe.push_device(raw.data_ptr(), n)
        pipe.pull_into(out_host.data_ptr(), out_host.numel())
    pipe.drain(out_host.data_ptr(), out_host.numel())
    launches0 = pipe.launches()
    barrier()
    t0 = time.perf_counter()
    pipe.span_begin()
    nb_dev, checked = 0, False
    for _ in range(args.steps):
        pipe.reset()
        pipe.push_device(raw.data_ptr(), n)
        if w["kind"] == "chain":
            nb = pipe.pull_into(out_host.data_ptr(), out_host.numel())  # CADUs of the batches decoded so far (keeps the output buffer drained)
            if nb and not checked and first is not None:
                checked = True
                frames_ok = frames_ok and np.array_equal(out_host[:nres * unit].numpy().reshape(-1, unit), res0.reshape(-1, unit))
            nb_dev += nb
        else:
            nb_dev += nres * unit  # (device-resident: the symbols stay in HBM)
    wall_dev = pipe.span_end() * 1e-3  # CUDA events
    wall_dev_host = time.perf_counter() - t0
    nb_dev += pipe.drain(out_host.data_ptr(), out_host.numel())
    frames_ok = frames_ok and nb_dev == args.steps * nres * unit
    launches1 = pipe.launches()
    barrier()
    # ---- e2e: host buffers; H2D of step i+1, demodulation of step i and decoding of step i-1 overlap; D2H of the result every step
    t0 = time.perf_counter()
    pipe.prefetch(host.data_ptr(), n)  # the first batch's copy is inside the timed region
    nb_e2e = 0
    for i in range(args.steps):
        pipe.reset()
        if i + 1 < args.steps:
            pipe.prefetch(host.data_ptr(), n)
        pipe.push_host(host.data_ptr(), n)
        nb_e2e += pipe.pull_into(out_host.data_ptr(), out_host.numel())
    nb_e2e += pipe.drain(out_host.data_ptr(), out_host.numel())
    torch.cuda.synchronize()
    wall_e2e = time.perf_counter() - t0
    frames_ok = frames_ok and nb_e2e == args.steps * nres * unit
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    pipe.set_pipelined(False)

    # the one exchange of this path: slowest rank's times, and every stream's counters on rank 0 (NCCL all-gather)
    wall_dev, wall_e2e = shard.max_over_ranks([wall_dev, wall_e2e], dev)
    all_ok = shard.all_true(frames_ok, dev)
    st = pipe.stats()
    fe = st.get("fec", {})
    counters = shard.gather_counters([n * args.steps, nres, fe.get("rs_corrected", 0), fe.get("rs_failed", 0), fe.get("replays", 0),
                                      st["demod"].get("repairs", 0), int(launches1 - launches0)], dev)
    if rank == 0:
        hbm, which = peaks()
        value = n * args.steps * world / wall_dev / 1e6
        e2e = n * args.steps * world / wall_e2e / 1e6
        sps = cfg.samplerate / cfg.symbolrate
        fir_ms = tim["agc_fir"]
        fir_bytes = bps + 8
        fir_roof = {"kernels": "k_agc_fir_w (convert + AGC + 31-tap RRC in one pass, one warp per range of tiles; the exact-seed launches return at once)", "bound": "hbm",
                    "achieved": n * fir_bytes / (fir_ms * 1e-3) / 1e9 if fir_ms > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                    "frac": (n * fir_bytes / (fir_ms * 1e-3) / 1e9 / hbm) if fir_ms > 0 else 0.0, "peak_source": which,
                    "traffic": measured_traffic("k_agc_fir_w", args.log2_samples) if cfg.fmt == "cs16" else None,
                    "note": f"{fir_bytes} B/sample ({cfg.fmt} in + cf32 out) over the event-timed FIR stage of a synchronous step"}
        line = {"metric": w["metric"], "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": wall_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32+u8" if w["kind"] == "chain" else "f32", "data": "synthetic",
                "config": {"workload": w["label"], "samples_per_step_per_gpu": n, "result_units_per_step_per_gpu": int(nres),
                           "result_unit_bytes": unit, "cadus_bit_exact_vs_transmitted": bool(all_ok) if w["kind"] == "chain" else None,
                           "cadus_with_parity_bytes_as_received": parity_as_received if w["kind"] == "chain" else None, "first_cadu_is_transmitted_frame": first,
                           "l2": "input batch (%.0f MiB) larger than L2" % (n * bps / 2 ** 20), "esn0_db": cfg.esn0_db},
                "e2e": {"value": e2e, "unit": "MS/s", "h2d_bytes_per_step": n * bps * world, "d2h_bytes_per_step": int(nres) * unit * world,
                        "pcie_h2d_GBps_alone": round(h2d_gbps, 2), "pcie_bound_MSps_per_gpu": round(h2d_gbps * 1e9 / bps / 1e6, 1),
                        "note": "%s is %d B/sample: the H2D copy of a step alone takes %.1f ms; e2e cannot exceed pcie_bound" % (cfg.fmt, bps, n * bps / h2d_gbps / 1e6)},
                "gpu_launches": int(launches1 - launches0), "host_wall_ms_per_step": wall_dev_host / args.steps * 1e3,
                "sync_mode": {"value": n * world / (sync_ms * 1e-3) / 1e6, "ms_per_step": sync_ms,
                              "note": "same step with every kernel alone on the GPU (decoder after the demodulator on the calling thread), rank 0"},
                "stage_ms_sync_step": {k: round(v, 4) for k, v in tim.items() if k != "vit_chunks"},
                "clocks": sampler.summary(), "stream_stats": st,
                "streams": [dict(zip(["samples", "result_units_per_step", "rs_corrected", "rs_failed", "replays", "demod_repairs", "gpu_launches"], c)) for c in counters]}
        if w["kind"] == "chain":
            chunk = 16384 if cfg.decoder == "metop" else max(cfg.cadu_bytes * 8, 8192)
            fbits = chunk * 3 // 4 if cfg.decoder == "metop" else chunk // 2
            alg = chunk + fbits // 8  # Viterbi of one chunk: int8 soft in + packed decoded bits out
            vit_ms = tim["k_vit_acs"]
            achieved = tim["vit_chunks"] * alg / (vit_ms * 1e-3) / 1e9 if vit_ms > 0 else 0.0
            line["mode"] = "pipelined chain: decoder one batch behind the demodulator (own stream + worker thread); K fresh streams + drain inside the timed region"
            line["roofline"] = {"kernel": "k_vit_acs3 (warp-per-chunk add-compare-select, the longest kernel of the step)", "bound": "hbm", "achieved": achieved,
                                "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": measured_traffic("k_vit_acs3", args.log2_samples),
                                "peak_source": which, "alg_bytes_per_chunk": alg,
                                "note": "timed alone (synchronous step); integer ACS kernel bound by the SM's ALU / shuffle-vote-reduce ports, not by HBM (DESIGN.md 6)"}
            line["roofline_fir_stage"] = fir_roof
        else:
            line["mode"] = "demodulator alone (no Costas loop, no decoder): K fresh streams, every push synchronous"
            out_b = 8.0 / sps
            line["roofline"] = fir_roof
            line["roofline_whole_step"] = {"bound": "hbm", "alg_bytes_per_sample": bps + out_b, "achieved": value * 1e6 * (bps + out_b) / 1e9 / world,
                                           "peak": hbm, "unit": "GB/s", "frac": value * 1e6 * (bps + out_b) / 1e9 / world / hbm,
                                           "note": "compulsory traffic of the step (%s in, cf32 symbols out) over the device-resident throughput of one GPU" % cfg.fmt}
        if world == 1 and not args.no_cpu_baseline:
            try:
                m = min(n, 1 << 25)
                hraw = host.numpy()[:(m if cfg.fmt == "cf32" else 2 * m)]
                cb, cpu_out = cpu_baseline(w, cfg, hraw, m, repeats=3, detail=False)
                line["cpu_baseline"] = cb
                if w["kind"] == "chain" and cpu_out is not None:  # the GPU's CADUs against the reference's on the common prefix
                    a, b = cpu_out.reshape(-1, unit), res0.reshape(-1, unit)
                    k = min(a.shape[0], b.shape[0])
                    line["config"]["cadus_identical_to_reference_on_prefix"] = {"frames_compared": int(k), "identical": bool(k > 0 and np.array_equal(a[:k], b[:k]))}
            except Exception as ex:  # the oracle is test infrastructure: its absence must not break the product bench
                line["cpu_baseline"] = {"value": None, "unit": "MS/s", "cores": 0, "kind": "unavailable", "sample": str(ex)[:200]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def pin_to_gpu_numa_node(index):
    """Run this rank (and allocate its pinned staging memory) on the CPUs of the GPU's NUMA node: with several ranks on one host the
    H2D copies otherwise cross the socket interconnect (round 1: 2 ranks reached 0.84 of 2x one rank end to end)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


if __name__ == "__main__":
    main()
