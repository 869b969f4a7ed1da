#!/usr/bin/env python
"""bench.py - beam=8 utterances/s of the chinese-asr inference hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU port of the reference path

A "step" is one pass of the whole hot path (log-mel features -> BiLSTM encoder -> bw=8 beam decode,
40 steps -> finalisation) over one batch of synthetic 10 s / 16 kHz utterances per GPU.  Workload =
BASELINE.json configs[4] (the one the metric is quoted on): 4096 utterances sharded over 8 GPUs =
512 utterances per GPU per step; per-GPU work is fixed as N grows (weak scaling), utterances are
independent so there is no data-path collective, only the final gather of hypotheses.
Prints ONE JSON line (rank 0)."""
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

SECONDS = 10.0
SR = 16000
MAX_LEN = 40
V = 5004


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=512,
                    help="utterances per GPU per step (configs[4]: 4096 utterances over 8 GPUs).  Measured: 592 "
                         "(4 attention CTAs per SM) gives +2.8 %% utterances/s, 444 gives -0.5 %%")
    ap.add_argument("--bw", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=SECONDS)
    ap.add_argument("--cpu-sample", type=int, default=64, help="utterances per batch of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rec-chunks", type=int, default=0,
                    help="recurrence chunks per direction (asr_set_recurrence_chunks, experiments); 0 = the default 7")
    ap.add_argument("--pipeline", type=int, default=2,
                    help="engines (handles) per GPU with batches in flight: the latency-bound encoder of one batch "
                         "overlaps the decoder of another (chinese_asr_b200.parallel.BatchPipeline); 1 = one handle")
    return ap.parse_args()


def synth_batch(B, n, seed):
    rng = np.random.default_rng(seed)
    return (0.1 * rng.standard_normal((B, n))).astype(np.float32)


def synth_batch_s16(B, n, seed):
    """The same waveforms as 16-bit PCM - what a WAV file holds and what the front end ships to the GPU."""
    return np.clip(np.round(synth_batch(B, n, seed) * 32768.0), -32768, 32767).astype(np.int16)


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
def cpu_port_rate(n_utts, bw, seconds, threads=None, reps=1, warm=0):
    """The oracle (CPU port of the reference path, same torch CPU library ops the reference calls)
    on a bounded sample: features + encoder + beam decode, batch = n_utts.  Returns utt/s."""
    import torch
    from oracle import asr_oracle as O
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use all the host threads it can
    torch.set_num_threads(threads or max(1, os.cpu_count() or 1))
    weights = O.make_weights(1234, "plain")
    n = int(seconds * SR)
    pcm = O.pcm_from_int16(synth_batch_s16(n_utts, n, 4242))     # fast_read (data.py:109-121) of 16-bit samples
    best = None
    for it in range(warm + reps):
        t0 = time.perf_counter()
        feats = [O.features(pcm[i]) for i in range(n_utts)]
        lens = torch.tensor([f.size(0) for f in feats])
        O.beam_decode(weights, bw, feats, lens, None)
        dt = time.perf_counter() - t0
        if it >= warm:
            best = dt if best is None else min(best, dt)
    return n_utts / best, best, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count()
    n_utts = args.cpu_sample
    times = []
    for it in range(args.warmup + args.steps):
        rate, dt, threads = cpu_port_rate(n_utts, args.bw, args.seconds)
        if it >= args.warmup:
            times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = n_utts / (ms / 1000.0)
    sample = f"{n_utts} utterances x {args.seconds:g} s per step (one batch), features+encoder+beam bw={args.bw}"
    line = {
        "impl": "reference", "metric": "utterances_per_sec_beam8", "value": value, "unit": "utt/s",
        "rtfx": value * args.seconds, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"bw={args.bw} beam decode of {args.seconds:g} s 16 kHz utterances "
                               f"(BASELINE.json configs[4]); CPU arm runs a bounded sample per step",
                   "beam": args.bw, "utt_seconds": args.seconds, "max_len": MAX_LEN,
                   "utts_per_step": n_utts, "cpu_baseline_sample": sample},
        "cpu_baseline": {"value": value, "unit": "utt/s", "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": cores},
        "e2e": {"value": value, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def decoder_step_bytes(B, k, L):
    """algorithmic HBM bytes of one decoder step (fp32 storage): decoder weights once,
    keys + encoder memory of every frame once per utterance (shared by its k beams), read + write of h / c / ctx per
    row, one embedding row per row.  The logits are NOT in this number: they no longer reach HBM (the vocabulary
    GEMM's epilogue keeps log-sum-exp partials and the top candidates per tile instead)."""
    w_dec = 4 * (2625536 + 65536 + 128 + 5129100)
    return w_dec + B * L * 2560 + B * k * 12288 + B * k * 1024


def run_native(args):
    import torch
    import torch.distributed as dist
    from oracle import asr_oracle as O          # weights / synthetic inputs only (not on the timed path)
    from chinese_asr_b200.model import Model
    from chinese_asr_b200.gpd import gpd
    from chinese_asr_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = None
    if world > 1:     # before any engine / NCCL thread exists: this rank's own cores next to its GPU
        cores = parallel.pin_rank_to_local_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        if cores:     # host-side torch work (weight initialisation) must not run more threads than this rank has cores
            torch.set_num_threads(max(1, min(torch.get_num_threads(), len(cores))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gpd["verbose"] = False
    gpd["temperature"] = 1.0
    B, k = args.batch, args.bw
    n = int(args.seconds * SR)
    L = (1 + (n - 1 - 512) // 160) // 3

    S = max(1, args.pipeline)
    weights = O.make_weights(1234, "plain")
    models = []
    for _ in range(S):
        mm = Model()
        mm.load_state(weights)
        mm.reserve(B, B * L, k, B * n, MAX_LEN)
        if args.rec_chunks:
            mm.set_recurrence_chunks(args.rec_chunks)
        models.append(mm)
    m = models[0]
    pipe = parallel.BatchPipeline(models)
    off = (np.arange(B + 1, dtype=np.int64) * n)
    # 16-bit PCM (what WAV files hold): converted to float32 by the log-mel kernel (asr_transcribe_pcm, ASR_PCM_S16).
    # End-to-end steps cycle through 2 pinned host buffers per engine.
    hosts = [torch.from_numpy(synth_batch_s16(B, n, 1000 + rank + 17 * i).reshape(-1)).pin_memory()
             for i in range(2 * S)]
    resident = hosts[0].to(dev)
    total = B * world

    gather_s = []        # host seconds of every end-of-run gather (pack + H2D + all_gather + D2H + scatter)

    def run_steps(e2e, steps, first=0):
        """`steps` batches of B utterances per GPU, handed round-robin to the S engines (each on its own host
        thread and stream).  e2e: every batch starts in pinned host memory (its H2D copy is inside the call) and
        ends with the hypotheses on the host; otherwise the PCM is resident in HBM."""
        if e2e:
            # a server's pipeline: every engine stages its next batch's PCM while it decodes the current one
            results = pipe.map([(hosts[(first + i) % (2 * S)], off) for i in range(steps)], prefetch=True, bw=k)
        else:
            results = pipe.map([(resident, off)] * steps, bw=k, resident=True)
        if world == 1:
            return results[-1]
        t_g = time.perf_counter()
        # the one collective of the path: a single gather of the hypotheses of every batch
        # this rank decoded, at the end of the run (utterance id = (step * world + rank) * B + i)
        rec = np.concatenate([parallel.pack_records(np.arange(B) + (s_ * world + rank) * B, tok, ln, sc, MAX_LEN)
                              for s_, (tok, ln, sc) in enumerate(results)])
        out = parallel.gather_hypotheses(rec, steps * total, MAX_LEN, device=dev)
        gather_s.append(time.perf_counter() - t_g)
        return out

    def timed(e2e, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run_steps(e2e, steps)          # returns when every batch's results are on the host
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    # warm-up: every engine alone first (workspaces, launch-time statics, graph capture on its second batch) ...
    pipe.each(lambda mm: [mm.transcribe(resident, off, bw=k, resident=True) for _ in range(2)])
    pipe.each(lambda mm: mm.transcribe(hosts[0], off, bw=k))
    # ... then W untimed steps through the pipeline
    run_steps(False, max(args.warmup, 3))
    run_steps(True, S)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for mm in models:
        mm.launch_count(reset=True)
    ms_dev = timed(False, args.steps)
    launches = sum(mm.launch_count(reset=True) for mm in models)
    ms_e2e = timed(True, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # instrumented repeat of the same step: per-stage CUDA-event durations on the launch stream
    # (one engine alone: the stage durations of a solo batch)
    m.stage_timing(True)
    pipe.each(lambda mm: mm.transcribe(resident, off, bw=k, resident=True) if mm is m else None)
    stages = m.stage_times()
    m.stage_timing(False)
    pipe.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    ms_step = ms_dev / args.steps
    value = total / (ms_step / 1000.0)
    e2e_val = total / (ms_e2e / args.steps / 1000.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1590.0))
    peak_src = "measured" if peaks else "fallback"

    # roofline of the dominant KERNEL.  The tcgen05 GEMM kernel (encoder input projections, attention
    # keys / queries, decoder LSTM cell, vocabulary projection) is the largest consumer of GPU time
    # (profiles/r01_launches.md); its launches are bracketed by CUDA events on the launch stream.
    gemm_ms = stages.pop("gemm_kernel_ms")
    split_ms = stages.pop("operand_split_ms")
    gemm_gflop = stages.pop("gemm_gflop")
    att_ms = stages.pop("attention_kernel_ms")
    R = B * k
    n_gemm = 4 + 1 + 3 * MAX_LEN
    roof_src = {}
    try:        # ncu DRAM bytes per GEMM-engine launch of this workload (tools/prof_cmd.sh -> profiles/r02f_roofline.json)
        roof_src = json.load(open(os.path.join(ROOT, "profiles", "r02f_roofline.json")))
    except Exception:
        pass
    kernel_ms = dict(stages)
    kernel_ms["gemm engine (all GEMM stages)"] = gemm_ms
    dom = max(("gemm engine (all GEMM stages)", "enc_recurrence", "attention", "topk_bookkeep", "features"),
              key=lambda kk: kernel_ms[kk])
    if dom.startswith("gemm"):
        ach = gemm_gflop / gemm_ms                        # GFLOP / ms = TFLOP/s (algorithmic fp32 2*M*N*K)
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the 125 GEMM-engine launches of one
        # pass of this workload (ncu, profiles/r02f_pass_metrics.csv; captured at 512 utterances per step - the
        # operand / output bytes scale with the rows)
        gb = roof_src.get("gemm_dram_bytes_per_pass")
        traffic = gb / n_gemm * B / 512 if gb and (k, L) == (8, 332) else None
        roof = {"kernel": "tc::gemm_split_pair_kernel", "bound": "tensor", "achieved": ach, "peak": tf_peak,
                "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": traffic,
                "peak_source": peak_src + " cuBLAS bf16 sustained.  Per 16 values of k the kernel issues one fp16 MMA "
                               "(a_hi*w_hi) and two bf16 MMAs (a_lo*w + a*w_lo over 2K): 3x the bf16 tensor-pipe time "
                               "per algorithmic FLOP, so tensor_pipe_frac = 3 * frac is the share of the pipe's peak",
                "tensor_pipe_frac": 3.0 * ach / tf_peak, "algorithmic_gflop_per_step": gemm_gflop,
                "ms_per_launch": gemm_ms / n_gemm, "launches_per_step": n_gemm, "ms_per_step": gemm_ms}
    elif dom == "enc_recurrence":
        fl = 2.0 * B * L * 4 * 2 * 1024 * 256
        ach = fl / (stages[dom] / 1000.0) / 1e12
        roof = {"kernel": "rec3::lstm_rec_tc3_kernel", "bound": "tensor", "achieved": ach, "peak": tf_peak,
                "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None, "peak_source": peak_src,
                "note": "latency-bound chain of 4*L dependent steps", "ms_per_launch": stages[dom] / 4,
                "launches_per_step": 4}
    else:
        if dom == "attention":
            byts = MAX_LEN * (B * L * 2560 + R * 3 * 2048 + 4 * (65536 + 128))
        elif dom == "features":
            byts = B * (4 * n + 4 * 720 * L + 2 * 4 * 80 * 3 * L)
        else:
            byts = MAX_LEN * (2 * 4 * R * V)
        ach = byts / (stages[dom] / 1000.0) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src}
    stages["gemm_kernel_ms"] = gemm_ms
    stages["operand_split_ms"] = split_ms
    stages["attention_kernel_ms"] = att_ms
    # the north-star group: one decoder step (cell + attention + projection + top-k) vs HBM roofline
    dec_ms = (stages["dec_cell"] + stages["attention"] + stages["vocab_proj"] + stages["topk_bookkeep"]) / MAX_LEN
    dec_bytes = decoder_step_bytes(B, k, L)
    # the memory-bound kernel of that group: attention scores + context over the encoder memory.  Per launch it
    # must read keys_exp (128 floats) + memory (512 floats) of every frame once per utterance, the queries, and
    # write ctx + its split operand
    att_bytes = B * L * 2560 + R * (4 * 128 + 4 * 512 + 6 * 512)
    att_us = 1000.0 * att_ms / MAX_LEN
    dec_roof = {"ms_per_decoder_step": dec_ms, "algorithmic_bytes": dec_bytes,
                "achieved_gbs": dec_bytes / (dec_ms / 1000.0) / 1e9,
                "frac_of_hbm_peak": dec_bytes / (dec_ms / 1000.0) / 1e9 / hbm_peak,
                "memory_bound_kernel": {"kernel": "attention_stream_kernel", "bound": "hbm", "us_per_launch": att_us,
                                        "algorithmic_bytes": att_bytes, "achieved": att_bytes / att_us / 1e3,
                                        "peak": hbm_peak, "unit": "GB/s", "frac": att_bytes / att_us / 1e3 / hbm_peak,
                                        "peak_source": peak_src},
                "kernels_per_step": ["cell GEMM (LSTM epilogue)", "query GEMM", "attention_stream_kernel",
                                     "vocabulary GEMM (LSE + top-k epilogue)", "beam_merge_kernel"],
                "note": "algorithmic bytes exclude logits (not materialised). The GEMMs of the step (cell, query, "
                        "vocabulary) are tensor-bound in fp32-faithful split precision (3 MMA slots per product), so "
                        "the whole step sits below the HBM roofline; the attention kernel is the memory-bound part"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N = 1 only
        rate, dt, threads = cpu_port_rate(args.cpu_sample, k, args.seconds, reps=3)
        cpu = {"value": rate, "unit": "utt/s", "cores": threads, "kind": "port", "host_cpus": os.cpu_count(),
               "sample": f"3 batches of {args.cpu_sample} utterances x {args.seconds:g} s, features+encoder+beam "
                         f"bw={k}, best batch {dt:.1f} s wall"}
    line = {
        "metric": "utterances_per_sec_beam8", "value": value, "unit": "utt/s", "rtfx": value * args.seconds,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"bw={k} beam decode of {args.seconds:g} s 16 kHz utterances, {B} per GPU per step "
                               f"(BASELINE.json configs[4]: 4096 utterances over 8 GPUs)",
                   "beam": k, "utts_per_gpu_per_step": B, "utt_seconds": args.seconds, "max_len": MAX_LEN,
                   "engines_per_gpu": S,
                   "pipeline": ("%d handles per GPU, batches handed round-robin, each engine on its own host thread and "
                                "stream: one batch's encoder overlaps another's decoder" % S) if S > 1 else "one handle",
                   "enc_frames_per_utt": L, "weights": "random-init (reference initialisers), fp32",
                   "cpu_baseline_sample": f"{args.cpu_sample} utterances per CPU batch (the GPU arm: {B} per step)",
                   "rank_cores": cores,
                   "final_gather_ms": [round(1e3 * g, 3) for g in gather_s[-2:]],
                   "pcm": "int16 (16-bit WAV samples), converted on the device",
                   "l2_policy": "inputs larger than L2 (PCM %.0f MB, gate pre-activations %.0f MB per step)"
                                % (B * n * 2 / 1e6, B * L * 8192 / 1e6)},
        "e2e": {"value": e2e_val, "unit": "utt/s", "rtfx": e2e_val * args.seconds,
                "h2d_bytes_per_step": int(B * n * 2 + (B + 1) * 16 + 3 * B * L * 4),
                "d2h_bytes_per_step": int(B * (MAX_LEN + 2) * 4 + 16),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "decoder_step_roofline": dec_roof,
        "stage_ms": stages,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if int(os.environ.get("RANK", "0")) == 0 and (args.impl == "reference" or world == 1):
        # torchrun exports OMP_NUM_THREADS=1; the CPU arms (the reference arm, and the cpu_baseline leg of the
        # native arm at N = 1) run on rank 0 and must see all host threads.  This has to happen before torch creates
        # its OpenMP pool (set_num_threads later is not enough).  NOT for the native arm at N > 1: there rank 0 pins
        # itself to its own slice of the cores like every other rank (parallel.pin_rank_to_local_cores), and an
        # OpenMP pool of os.cpu_count() spinning threads on that slice turns the host-side weight initialisation
        # into minutes (observed: the 8-rank run did not finish in 10 minutes).
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        os.environ["MKL_NUM_THREADS"] = str(os.cpu_count() or 1)
    if args.impl == "native":
        # the whole native run takes 1 - 3 minutes: a stuck one dumps every thread's stack and exits instead of
        # hanging until the caller's timeout
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("ASR_BENCH_WATCHDOG_S", "600")), exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
