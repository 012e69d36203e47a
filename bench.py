#!/usr/bin/env python
"""Headline benchmark: autoregressive mel frames/s at batch 32 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host CPU

One "step" = one full synthesis of the workload: encoder, 1000 K/V-cached decode steps at
B=32 / S=258 (stop disabled so every sample emits all frames), Postnet.  `value` is measured with
the inputs already resident in HBM; `e2e` is the same job through the public API from pinned
host buffers with the results copied back.  Multi-GPU = independent replicas (decode does not
shard across GPUs: SURVEY.md §8e), one process per GPU, max time over ranks.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "mel frames/sec autoregressive @ batch 32"
UNIT = "frames/s"


# ------------------------------------------------------------------------------------------------
# workload + algorithmic bytes (DESIGN.md "Roofline")
# ------------------------------------------------------------------------------------------------
def decode_step_params(cfg):
    """Decoder parameters streamed on every step: all of decoder.* except the six kv_transforms."""
    D, F, P, M, L = cfg.decoder_hidden, 4 * cfg.decoder_hidden, cfg.prenet_hidden, cfg.num_mels, cfg.n_decoder_layer
    per_layer = 3 * D * D + D * D + D * D + D * D + 2 * D * F + 6 * D
    return P * M + P + P * P + P + D * P + 1 + L * per_layer + 2 * D + M * D + D + 1


def decode_bytes(cfg, batch, mem_len, n_steps, elem=4):
    """Algorithmic HBM bytes of n_steps decode steps (weights once per step, cross K/V once per step,
    self K/V of all previous frames, the appended K/V row, frame in/out)."""
    kv = 2 * cfg.n_decoder_layer * cfg.decoder_hidden
    w = decode_step_params(cfg) * elem
    total = 0
    for t in range(n_steps):
        total += w + batch * kv * elem * mem_len + batch * kv * elem * (t + 1) + batch * kv * elem \
            + batch * cfg.num_mels * 4 * 2
    return total


KERNEL_NAME = {4: "pipelined_decode_kernel"}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the decode kernel (50 steps, t = 475..524, B=32,
    S=258), read from the committed `ncu --set full` raw page of the shipped build (profiles/README.md) - never a
    literal in this file.  Returns (bytes or None, source)."""
    import csv
    for name in ("r2_pipe_ncu_raw.csv", "r1_pipe_ncu_raw.csv"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        try:
            rows = list(csv.reader(open(path)))
            head = rows[0]
            rd, wr = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
            units, vals = rows[1], rows[2]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            total = sum(float(vals[i].replace(",", "")) * scale[units[i]] for i in (rd, wr))
            return total, "profiles/" + name
        except (ValueError, IndexError, KeyError):
            continue
    return None, None


def decode_bytes_range(cfg, batch, mem_len, t0, t1, elem=4):
    return decode_bytes(cfg, batch, mem_len, t1, elem) - decode_bytes(cfg, batch, mem_len, t0, elem)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is read in-process
    from a thread (nvidia_ml_py); the `nvidia-smi -lms` subprocess of the recipe is the fallback - it costs ~4 % of
    this job's wall clock (20 cooperative launches per synthesis stall behind its queries), NVML reads do not."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period=0.05):
        self.rows, self.proc, self.index, self.period = [], None, index, period
        self.stop, self.thread, self.source = threading.Event(), None, None

    def _nvml_loop(self, nv, h):
        names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                self.rows.append([str(sm), str(mx), "%.1f" % pw] + ["Active" if bits & b else "Not Active" for _, b in names])
            except Exception:
                pass
            self.stop.wait(self.period)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        self.stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0, "source": self.source}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "source": self.source}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reduce_max_over_ranks(value, world):
    """max over ranks of a python float (the slowest rank defines the job's time)."""
    if world == 1:
        return value
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, seconds_this_rank, world):
    """Whole-job throughput of `world` replicas: all units / slowest rank's time."""
    return units_per_rank * world / reduce_max_over_ranks(seconds_this_rank, world)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference algorithm (uncached O(T^2) loop) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(batch_size, text_len, horizon, repeats=1):
    """The oracle's faithful restatement of synthesize.eval_batch (re-runs the decoder over all
    frames so far each step), B x text_len, truncated to `horizon` frames.  Returns frames/s."""
    from oracle import tts_oracle as O
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    cfg = O.ModelConfig()
    params = O.synth_params(cfg, seed=0)
    params["decoder.stop_net.bias"] = torch.tensor([-1e4])
    batch = O.synth_batch(cfg, batch=batch_size, text_len=text_len, n_frames=4, seed=1)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            out = O.eval_batch_uncached(params, cfg, batch, max_frames=horizon)
            times.append(time.perf_counter() - t0)
    frames = out["mel_pre"].shape[0] * out["mel_pre"].shape[1]
    return frames / min(times), times, frames


def run_vocoder(args, dev):
    """utils/audio.py:60-79 (mel2wav, one utterance per CPU worker in synthesize.py:82,99) as ONE batched GPU call of
    tts_b200.vocoder on the shape the synthesis produces (B utterances x `frames` mel frames, synthetic mels in the model's range)."""
    from tts_b200 import vocoder as V
    g = torch.Generator().manual_seed(7)
    mels = (torch.randn(args.batch, args.frames, 80, generator=g) * 0.6 + torch.linspace(1.5, -2.5, 80)).clamp_(-4, 4).to(dev)
    eng = V.GriffinLim(dev)
    lens = [args.frames] * args.batch
    for _ in range(2):
        eng(mels, lens)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng(mels, lens)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 3
    audio_s = args.batch * 200 * (args.frames - 1) / 16000.0
    return {"stage": "mel2wav: Griffin-Lim, 60 iterations, n_fft 2048 / hop 200 / win 800 (hyperparams.py:7-18)", "ms": ms,
            "frames_per_s": args.batch * args.frames / (ms / 1e3), "audio_seconds": audio_s,
            "realtime_factor": audio_s / (ms / 1e3), "gpu_launches": 2 * 61 + 2}


def run_module_api_loop(args, dev, params, dev_batch):
    """frames/s of the reference's own loop shape (synthesize.py:35-56) driving transformer.tacotron.Tacotron: encoder once,
    then per frame torch.cat + Decoder.forward(leave_one=True) + stop bookkeeping + `torch.all(finished)` (a host sync)."""
    from tts_b200.config import hparams_from
    from tts_b200 import synthetic as O
    from transformer import tacotron
    cfg = O.ModelConfig(max_generation_frames=args.frames)
    hp = hparams_from(cfg)
    m = tacotron.Tacotron(hp)
    m.load_state_dict(params, strict=True)
    m.to(dev).eval()
    horizon = min(args.frames, args.module_api_frames)
    n = dev_batch["inputs"].shape[0]

    def loop():
        with torch.no_grad():
            lengths = torch.ones([n], dtype=torch.int32, device=dev)
            finished = torch.zeros([n], dtype=torch.bool, device=dev)
            mels = torch.zeros([n, 0, cfg.num_mels], dtype=torch.float32, device=dev)
            enc = m.encoder(dev_batch["inputs"], dev_batch["input_lengths"], dev_batch["input_spk_ids"], dev_batch["input_language_vecs"])
            while not torch.all(finished) and mels.shape[1] < horizon:
                dec_in = torch.cat([mels, torch.zeros([n, 1, cfg.num_mels], device=dev)], dim=1)
                mel_bef, stop_logits, _ = m.decoder(enc, dev_batch["input_lengths"], dec_in, lengths, leave_one=True)
                stop = stop_logits[:, -1] > 0
                mels = torch.cat([mels, mel_bef[:, -1:]], dim=1)
                finished = torch.logical_or(finished, stop)
                lengths = torch.where(finished, lengths, lengths + 1)
            return mels + m.postnet(mels, lengths)

    loop()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    out = loop()
    torch.cuda.synchronize(dev)
    secs = time.perf_counter() - t0
    return {"value": out.shape[0] * out.shape[1] / secs, "unit": UNIT, "frames": int(out.shape[1]),
            "us_per_frame_step": 1e6 * secs / out.shape[1],
            "note": "reference loop shape (synthesize.py:35-56) over transformer.tacotron.Tacotron: one cooperative launch, two "
                    "torch.cat and one host sync per frame; first %d of %d frames (the torch.cat cost grows with t)" % (horizon, args.frames)}


def gpu_eager_sample(dev, batch_size, text_len, horizon, frames):
    """the K/V-cached algorithm as plain PyTorch eager on the SAME B200 (cuBLAS fp32, TF32 off) -
    the de-facto existing Blackwell path for a K/V-cached decode, and the honest bar for our kernel.  This is the
    oracle's eval_batch_cached executed on cuda; it is a reported baseline like cpu_baseline, never the product."""
    from oracle import tts_oracle as O
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        cfg = O.ModelConfig()
        params = {k: v.to(dev) for k, v in O.synth_params(cfg, seed=0).items()}
        params["decoder.stop_net.bias"] = torch.tensor([-1e4], device=dev)
        batch = {k: (v.to(dev) if torch.is_tensor(v) else v)
                 for k, v in O.synth_batch(cfg, batch=batch_size, text_len=text_len, n_frames=4, seed=1).items()}
        with torch.no_grad():
            O.eval_batch_cached(params, cfg, batch, 8)     # warm-up (cuBLAS handles, allocator)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            out = O.eval_batch_cached(params, cfg, batch, horizon)
            torch.cuda.synchronize(dev)
            secs = time.perf_counter() - t0
        n = out["mel_pre"].shape[0] * out["mel_pre"].shape[1]
        return {"value": n / secs, "unit": UNIT, "kind": "torch eager (cuBLAS fp32, TF32 off), K/V-cached oracle loop on cuda",
                "sample": "B=%d S=%d, first %d of %d frames (per-step cost grows with t), %.1f s"
                          % (batch_size, text_len, horizon, frames, secs)}
    except Exception as exc:   # a baseline must never take the benchmark down
        return {"value": None, "unit": UNIT, "error": repr(exc)[:200]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    threads = torch.get_num_threads()
    horizon = args.ref_horizon
    steps_fps = []
    for i in range(args.warmup + args.steps):
        fps, times, frames = cpu_reference_sample(args.batch, args.text_len, horizon)
        if i >= args.warmup:
            steps_fps.append((fps, times[0]))
    secs = sum(t for _, t in steps_fps)
    value = frames * len(steps_fps) / secs
    sample = ("oracle port of synthesize.eval_batch (uncached O(T^2) loop, transformer/*.py fp32) on host CPU, "
              "B=%d S=%d truncated to the first %d frames of %d (frames/s falls further with the horizon)"
              % (args.batch, args.text_len, horizon, args.frames))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(steps_fps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(args):
    return {"workload": "autoregressive synthesize.py decode, batch=%d, %d-byte text (S=%d tokens), %d frames, "
                        "1xB200 per replica (BASELINE.json configs[1])" % (args.batch, args.text_len - 2, args.text_len,
                                                                           args.frames),
            "batch": args.batch, "text_tokens": args.text_len, "frames": args.frames, "stop": "disabled (bias -1e4)",
            "storage": "fp32 weights + fp32 K/V cache", "parallelism": "replicas x%d (decode does not shard)" % args.gpus,
            "l2": "per-step working set (200 MB weights + >=300 MB K/V) exceeds the 126 MB L2; no explicit flush",
            "alignments": "encdec attention rows recorded on device (not copied to host)"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    from tts_b200 import synthetic as O     # workload definition: config, seeded weights and inputs (no oracle here)
    from tts_b200 import _native
    from tts_b200.engine import TtsEngine

    cfg = O.ModelConfig(max_generation_frames=args.frames)
    params = O.synth_params(cfg, seed=0)
    params["decoder.stop_net.bias"] = torch.tensor([-1e4])
    eng = TtsEngine.from_state_dict(params, cfg, dev)
    host_batch = O.synth_batch(cfg, batch=args.batch, text_len=args.text_len, n_frames=4, seed=1 + rank)
    keys = ("inputs", "input_lengths", "input_spk_ids", "input_language_vecs")
    pinned = {k: host_batch[k].pin_memory() for k in keys}
    dev_batch = {k: v.to(dev) for k, v in pinned.items()}
    sess = eng.new_session(args.batch, args.text_len, args.frames, "encdec")
    lib = _native.load()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def job_resident():
        return eng.generate(dev_batch, max_frames=args.frames, record_align="encdec", chunk=args.chunk,
                            impl=args.decode_impl, session=sess)

    out_host = {}

    def job_e2e():
        b = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        out = eng.generate(b, max_frames=args.frames, record_align="encdec", chunk=args.chunk,
                           impl=args.decode_impl, session=sess)
        for k in ("mel_pre", "mel_aft", "generated_lengths"):
            if k not in out_host:
                out_host[k] = torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
            out_host[k].copy_(out[k], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out

    def timed(job, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = job()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / 1e3, out

    for _ in range(max(args.warmup, 3)):
        job_resident()
    lib.tts_launch_count_reset()
    with ClockSampler(local) as clocks:
        secs, out = timed(job_resident, args.steps)
    launches = int(lib.tts_launch_count())
    frames_per_job = int(out["mel_pre"].shape[0] * out["mel_pre"].shape[1])
    assert frames_per_job == args.batch * args.frames, "stop fired unexpectedly"
    value = aggregate_throughput(frames_per_job * args.steps, secs, world)

    job_e2e()
    secs_e2e, out = timed(job_e2e, args.steps)
    e2e_value = aggregate_throughput(frames_per_job * args.steps, secs_e2e, world)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())

    # ---- roofline of the dominant kernel: the decode steps, timed alone with CUDA events ----------
    mem = eng.encode(dev_batch["inputs"], dev_batch["input_lengths"], dev_batch["input_spk_ids"],
                     dev_batch["input_language_vecs"])
    dec_times = []
    for _ in range(3):
        sess.begin(mem, dev_batch["input_lengths"])
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done = 0
        while done < args.frames:
            n = min(args.chunk, args.frames - done)
            sess.step(n, update_state=True, impl=args.decode_impl)
            done += n
        e1.record()
        torch.cuda.synchronize(dev)
        dec_times.append(e0.elapsed_time(e1) / 1e3)
    dec_secs = min(dec_times)
    peak, peak_src = measured_peaks()
    alg_bytes = decode_bytes(cfg, args.batch, args.text_len, args.frames)
    achieved = alg_bytes / dec_secs / 1e9
    # one launch of the persistent kernel = `chunk` decode steps; report the average launch
    impl = args.decode_impl if args.decode_impl != 0 else 4      # the library's default is the pipelined kernel
    persistent = impl == 4
    n_launch = max(1, -(-args.frames // args.chunk)) if persistent else args.frames
    std_shape = persistent and args.chunk == 50 and args.frames == 1000 and args.batch == 32 and args.text_len == 258
    roofline = {"bound": "hbm", "kernel": "%s (one launch = %d decode steps; %d launches per job)"
                % (KERNEL_NAME[impl], args.chunk, n_launch) if persistent else "decode step (per-phase kernels)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic()[0] if std_shape else None,
                "traffic_note": "dram__bytes_read+write of the launch covering t=475..524, read from %s; "
                                "algorithmic bytes of that launch: %.3e" % (ncu_traffic()[1], decode_bytes_range(cfg, args.batch, args.text_len, 475, 525)),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / n_launch,
                "launch_seconds": dec_secs / n_launch, "us_per_decode_step": 1e6 * dec_secs / args.frames,
                "share_of_job": dec_secs / (secs / args.steps)}

    # ---- the UNCHANGED synthesize.py call pattern (one Decoder.forward per frame through the drop-in module API, a
    #      D2H sync per frame for `torch.all(finished)`), next to generate(): what a user who just swaps the package gets
    module_api = None
    if not args.no_module_api:
        try:
            module_api = run_module_api_loop(args, dev, params, dev_batch)
        except Exception as exc:
            module_api = {"error": repr(exc)[:300]}
    # ---- the stage after the path: mel -> waveform, 60 Griffin-Lim iterations on the synthesised batch shape
    vocoder = None
    if not args.no_vocoder and rank == 0:
        try:
            vocoder = run_vocoder(args, dev)
        except Exception as exc:
            vocoder = {"error": repr(exc)[:300]}
    train = None
    if not args.no_train:
        try:
            del sess, eng
            torch.cuda.empty_cache()
            train = run_train_step(args, dev, rank, world)
        except Exception as exc:   # the headline line must survive a failure of the secondary workload
            import traceback
            train = {"error": repr(exc)[:300], "trace": traceback.format_exc()[-600:]}
    line = None
    if rank == 0:
        # Baselines are timed at N=1 only: under torchrun the other ranks would spin in an NCCL barrier for the
        # whole CPU sample and their polling threads starve it (the round-1 lines at N>1 showed exactly that).
        cpu = eager = None
        if not args.no_cpu_baseline and world == 1:
            fps, times, frames = cpu_reference_sample(args.batch, args.text_len, args.ref_horizon)
            cpu = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "oracle port of synthesize.eval_batch (uncached O(T^2) loop) on host CPU, B=%d S=%d, "
                             "first %d of %d frames, %.1f s" % (args.batch, args.text_len, args.ref_horizon,
                                                                args.frames, times[0])}
            eager = gpu_eager_sample(dev, args.batch, args.text_len, args.eager_horizon, args.frames)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * reduce_max_over_ranks(secs, 1) / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args), "clocks": clocks.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "gpu_eager_baseline": eager, "decode_impl": args.decode_impl, "module_api_loop": module_api,
                "train_step": train, "mel2wav": vocoder}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0



def train_flops(S, T, B):
    """Algorithmic FLOPs of one teacher-forced training step per GPU: forward per SURVEY.md §8d (dense contractions +
    attention with causal skipping credited + Postnet), backward = 2 x forward."""
    fwd = B * (S * (37.75e6 + 4 * S * 512 * 6 + 14.16e6) + T * (99.09e6 + 4 * T * 768 * 6 * 0.5 + 4 * S * 768 * 6 + 0.565e6
                                                                + 0.124e6 + 8.68e6))
    return 3.0 * fwd


def gpu_eager_train_sample(dev, B, S, T, autocast_bf16):
    """SURVEY.md §8d(iii) for metric 2: the reference's training step as plain PyTorch eager on the SAME B200 - the oracle's
    functional restatement of Tacotron.forward (batch-statistics BatchNorm; dropout off, which only helps it) + compute_loss
    under torch autograd + torch.optim.Adam, cuBLAS / cuDNN kernels, materialised [B,H,T,T] attention as in the reference.
    fp32 with TF32 off is the reference's own arithmetic; bf16 autocast is what a user would switch on first.  A reported
    baseline like cpu_baseline, never the product."""
    from oracle import tts_oracle as O
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        cfg = O.ModelConfig()
        params = {k: v.to(dev) for k, v in O.synth_params(cfg, seed=0).items()}
        leaves = [v for k, v in params.items() if v.dtype.is_floating_point and "running_" not in k and "num_batches" not in k]
        for v in leaves:
            v.requires_grad_()
        opt = torch.optim.Adam(leaves, lr=1e-4, eps=1e-6)
        batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in O.synth_batch(cfg, batch=B, text_len=S, n_frames=T, seed=100).items()}

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast_bf16):
                out = O.tacotron_forward(params, cfg, batch, batch_stats=True)
                out = {k: (v.float() if torch.is_tensor(v) else v) for k, v in out.items()}
            loss = O.compute_loss(params, cfg, batch["mel_targets"], batch["target_lengths"], out)["loss"]
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss

        step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 2
        for _ in range(n):
            loss = step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / n
        return {"value": ms, "unit": "ms", "loss": float(loss.detach()), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                "kind": "torch eager on the same B200: oracle forward + compute_loss + autograd + torch.optim.Adam, %s, dropout off"
                        % ("bf16 autocast" if autocast_bf16 else "fp32, TF32 off")}
    except Exception as exc:   # a baseline must never take the benchmark down
        return {"value": None, "unit": "ms", "error": repr(exc)[:300]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.empty_cache()


def run_train_step(args, dev, rank, world):
    """BASELINE.json metric 2 / configs[2]: the teacher-forced TRAINING step (train.py:165-191: forward, compute_loss,
    backward, gradient all-reduce, Adam) at per-GPU batch 64 x 1000 mel frames, 258 text tokens, bf16 tensor-core
    kernels with fp32 master weights, dropout ON, through the drop-in transformer/ API.  Data parallel: one process
    per GPU, the 83.5 M gradients (333.9 MB fp32) all-reduced over NCCL every step (weak scaling: 64 per GPU)."""
    import torch.distributed as dist
    from tts_b200 import synthetic as O
    from tts_b200 import _native
    from tts_b200.config import hparams_from
    from tts_b200.dist import GradBuckets
    from tts_b200.optim import FusedAdam, l2_selected_names
    from transformer import tacotron
    cfg = O.ModelConfig()
    hp = hparams_from(cfg)
    hp.l2_in_optimizer = True
    torch.manual_seed(0)
    m = tacotron.Tacotron(hp)
    m.load_state_dict(O.synth_params(cfg, seed=0), strict=True)
    m.to(dev).train()
    B, S, T = args.tf_batch, args.text_len, args.frames
    host = O.synth_batch(cfg, batch=B, text_len=S, n_frames=T, seed=100 + rank)
    cfg4 = getattr(args, "cfg4", False)
    if cfg4:
        # BASELINE configs[3]: 38-language / 128-speaker mixed RAGGED batch, global 128 = 16 per GPU at 8 GPUs: text lengths
        # U[32, 258], mel lengths U[240, 800], padded to the longest (dataloader.py:419-439), one language / speaker per row
        B, S, T = 16, 258, 800
        g = torch.Generator().manual_seed(400 + rank)
        host = O.synth_batch(cfg, batch=B, text_len=S, n_frames=T, seed=100 + rank)
        in_len = torch.randint(32, S + 1, (B,), generator=g)
        tg_len = torch.randint(240, T + 1, (B,), generator=g)
        in_len[0], in_len[1], tg_len[0], tg_len[1] = 32, S, T, 240
        pos = torch.arange(S)[None, :]
        ids = host["inputs"]
        ids = torch.where(pos == (in_len[:, None] - 1), torch.ones_like(ids), ids)
        host["inputs"] = torch.where(pos < in_len[:, None], ids, torch.zeros_like(ids))
        host["mel_targets"] = (host["mel_targets"] * (torch.arange(T)[None, :, None] < tg_len[:, None, None])).contiguous()
        host["input_lengths"], host["target_lengths"] = in_len, tg_len
        host["input_spk_ids"] = (torch.arange(B) * 37 + 5 + 16 * rank) % 128
        lang = torch.zeros(B, cfg.max_num_language)
        lang[torch.arange(B), (torch.arange(B) * 7 + rank) % 38] = 1.0
        host["input_language_vecs"] = lang
    keys = ("inputs", "input_lengths", "mel_targets", "target_lengths", "input_spk_ids", "input_language_vecs")
    # the feeder's side of the boundary (dataloader.py:419-439): pageable host arrays, the language as an id; the stager
    # (tts_b200/staging.py) moves them through pinned arenas on a copy stream and expands the one-hot on the device
    from tts_b200.staging import BatchStager
    host_batch = {k: host[k] for k in keys if k != "input_language_vecs"}
    host_batch["input_language_ids"] = host["input_language_vecs"].argmax(dim=1)
    stager = BatchStager(dev, depth=2, n_languages=host["input_language_vecs"].shape[1])
    staged = [stager.stage(host_batch)]
    sel = l2_selected_names(m)
    opt = FusedAdam(m.parameters(), lr=hp.max_lr, eps=hp.adam_eps, reg_weight=hp.reg_weight,
                    l2_params=[p for n, p in m.named_parameters() if n in sel])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: tacotron.learning_rate_schedule(s, hp))
    ddp = None
    buckets = None
    if world > 1:
        if args.dp == "ddp":
            ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[dev.index], output_device=dev.index)
        else:
            # gradients appear per sub-module, in backward order: each group's all-reduce starts behind the rest of backward
            buckets = GradBuckets([list(m.postnet.parameters()), list(m.decoder.parameters()), list(m.encoder.parameters())])
            buckets.broadcast_parameters(0)
            if args.dp == "buckets":      # "buckets-late": everything is reduced after backward (no overlap), for comparison
                buckets.attach_hooks()
    fwd = ddp if ddp is not None else m
    lib = _native.load()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    ar_ms = []

    def step(timed_ar=False):
        batch = staged[0].wait()                 # dict_send_to (utils/__init__.py:3), staged one step ahead
        staged[0] = stager.stage(host_batch)     # the next batch's copy overlaps this step
        out = fwd(**batch)
        losses = tacotron.compute_loss(m, batch["mel_targets"], batch["target_lengths"], out, hp)
        opt.zero_grad()                       # train.py:173 (torch default: gradients set to None, assigned by backward)
        losses["loss"].backward()
        if buckets is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            buckets.finish()      # the hooked groups are already in flight: this is the EXPOSED part of the all-reduce
            e1.record()
            if timed_ar:
                ar_ms.append((e0, e1))
        opt.step()
        sched.step()
        loss_host.copy_(losses["loss"].detach(), non_blocking=True)
        batch.release()
        return losses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    lib.tts_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses = step(timed_ar=True)
    e1.record()
    barrier()
    ms = reduce_max_over_ranks(e0.elapsed_time(e1) / args.steps, world)
    launches = int(lib.tts_launch_count()) // max(args.steps, 1)
    # host side of one step: Python + launch time with an empty GPU queue (not part of the timed region)
    t_h0 = time.perf_counter()
    step()
    host_issue_ms = 1e3 * (time.perf_counter() - t_h0)
    barrier()
    eager = None
    if world == 1 and not args.no_cpu_baseline and not cfg4:
        torch.cuda.empty_cache()
        eager = {"bf16_autocast": gpu_eager_train_sample(dev, B, S, T, True), "fp32": gpu_eager_train_sample(dev, B, S, T, False)}
    flops = train_flops(S, T, B)
    valid_frames = B * T
    if cfg4:   # algorithmic FLOPs of the VALID tokens / frames (the padded rows a kernel also touches earn nothing)
        flops = sum(train_flops(int(s_), int(t_), 1) for s_, t_ in zip(host["input_lengths"], host["target_lengths"]))
        valid_frames = int(host["target_lengths"].sum())
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    ar = [a.elapsed_time(b) for a, b in ar_ms]
    n_params = sum(p.numel() for p in m.parameters())
    return {"metric": "teacher-forced train step ms", "value": ms, "unit": "ms", "higher_is_better": False, "n_gpus": world,
            "steps": args.steps, "scaling": "weak", "dtype": "bf16 (fp32 accumulate, fp32 master weights and optimizer state)",
            "config": {"workload": ("multilingual mixed ragged train step, per-GPU batch=%d (38 languages, 128 speakers, text U[32,%d], "
                                    "mel U[240,%d], padded to the longest), dropout on (BASELINE.json configs[3])" % (B, S, T)) if cfg4 else
                                   ("teacher-forced train step, per-GPU batch=%d x %d mel frames, %d text tokens, dropout on "
                                    "(BASELINE.json configs[2])" % (B, T, S)), "global_batch": B * world,
                       "parallelism": "dp%d (%s)" % (world, "single GPU" if world == 1 else
                                                     ("DistributedDataParallel" if ddp is not None else
                                                      "bucketed NCCL all-reduce overlapped with backward, 64 MB buckets"))},
            "frames_per_s": valid_frames * world / (ms / 1e3), "loss": float(loss_host),
            "h2d_bytes_per_step": stager.h2d_bytes // (max(args.warmup, 3) + args.steps + 1), "d2h_bytes_per_step": 4,
            "gpu_launches_per_step": launches, "host_issue_ms_per_step": host_issue_ms,
            "roofline": {"bound": "tensor", "achieved": flops / (ms / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (ms / 1e3) / 1e12 / peak, "traffic": None,
                         "flops_per_step_per_gpu": flops, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
            "gpu_eager_baseline": eager,
            "allreduce": None if not ar else {"bytes": 4 * n_params, "ms_median": statistics.median(ar),
                                              "note": "exposed time after backward (CUDA events around finish()); the postnet and decoder groups start from "
                                                      "autograd hooks and run behind the rest of backward"}}


def run_forward(args):
    """Secondary workload (not the headline): the teacher-forced FORWARD pass of BASELINE.json configs[2]
    (batch 64 x 1000 mel frames, 256-byte text) through TtsEngine.forward.  The backward pass and the NCCL gradient
    all-reduce of the training step are not built, so this is NOT the train-step metric."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    from tts_b200 import synthetic as O
    from tts_b200 import _native
    from tts_b200.engine import TtsEngine
    cfg = O.ModelConfig()
    eng = TtsEngine.from_state_dict(O.synth_params(cfg, seed=0), cfg, dev)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v)
             for k, v in O.synth_batch(cfg, batch=args.tf_batch, text_len=args.text_len, n_frames=args.frames, seed=1).items()}
    lib = _native.load()
    for _ in range(max(args.warmup, 3)):
        eng.forward(batch, want_align=False)
    torch.cuda.synchronize(dev)
    lib.tts_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = eng.forward(batch, want_align=False)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / args.steps
    # forward FLOPs per SURVEY.md §8d (dense contractions + attention + postnet), 2 flops per MAC
    S, T, B = args.text_len, args.frames, args.tf_batch
    flops = B * (S * (37.75e6 + 4 * S * 512 * 6 + 14.16e6) + T * (99.09e6 + 4 * T * 768 * 6 * 0.5 + 4 * S * 768 * 6 + 0.565e6
                                                                 + 0.124e6 + 8.68e6))
    print(json.dumps({"metric": "teacher-forced FORWARD ms (forward only: backward / all-reduce not built)", "value": ms,
                      "unit": "ms", "n_gpus": 1, "steps": args.steps, "higher_is_better": False, "dtype": "f32 (3xTF32 tcgen05 GEMMs)",
                      "data": "synthetic", "config": {"workload": "teacher-forced forward, batch=%d x %d mel frames, %d text tokens"
                                                       % (B, T, S)},
                      "gpu_launches": int(lib.tts_launch_count()) // args.steps,
                      "achieved_tflops_fp32_equivalent": flops / ms / 1e9,
                      "mel_bef_abs_mean": float(out["mel_bef"].abs().mean())}))
    return 0


def run_train_only(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    with ClockSampler(local) as clocks:
        line = run_train_step(args, dev, rank, world)
    line["clocks"] = clocks.summary()
    line["data"] = "synthetic"
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--text-len", type=int, default=258)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--chunk", type=int, default=50, help="decode steps per launch / host poll")
    ap.add_argument("--decode-impl", type=int, default=0,
                    help="0 default (= 4), 1 per-phase kernels, 2 CUDA graph, 4 pipelined kernel")
    ap.add_argument("--ref-horizon", type=int, default=96, help="frames of the CPU reference sample")
    ap.add_argument("--eager-horizon", type=int, default=300, help="frames of the torch-eager-on-GPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="decode", choices=["decode", "forward", "train"],
                    help="decode = the headline (default, also reports train_step); train = only the teacher-forced "
                         "training step (BASELINE metric 2); forward = teacher-forced forward pass only")
    ap.add_argument("--no-train", action="store_true", help="skip the train_step measurement of the default run")
    ap.add_argument("--no-module-api", action="store_true", help="skip the per-frame module-API loop measurement")
    ap.add_argument("--no-vocoder", action="store_true", help="skip the mel -> waveform (Griffin-Lim) measurement")
    ap.add_argument("--module-api-frames", type=int, default=300)
    ap.add_argument("--dp", default="buckets", choices=["buckets", "buckets-late", "ddp"], help="gradient exchange of the train step at N > 1")
    ap.add_argument("--tf-batch", type=int, default=64, help="batch of the --workload forward run")
    ap.add_argument("--cfg4", action="store_true", help="--workload train on the BASELINE configs[3] share: ragged mixed-language batch of 16 per GPU")
    args = ap.parse_args()
    if args.workload == "forward":
        return run_forward(args)
    if args.workload == "train":
        return run_train_only(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
