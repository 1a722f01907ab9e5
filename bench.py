"""bench.py -- headline benchmark of the B200-native tensorflow/compression hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    (N > 1: launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`)

Workload (BASELINE.json configs[1], "cfg2"): bls2017 compress path at batch 256, 256x256x3 images,
num_filters=128  ->  latents y [256,16,16,128] fp32 (256 code streams x 32768 symbols, 128 channel
tables, precision 12, overflow/Elias-gamma escape enabled).  One STEP = one pass of the entropy-bottleneck
hot path over one batch: quantise (y - offset -> rint -> - cdf_offset), range-encode every stream,
finalize and pack the strings (ContinuousBatchedEntropyModel.compress).  The GDN layers of the analysis
transform ([256,64,64,128] and [256,32,32,128]) and the decode path are measured in the same run and
reported in the `gdn` / `decode` objects of the JSON line.

`value`   = symbols / s with y resident in HBM (whole job, all ranks).
`e2e`     = the same metric through the public API with HOST buffers: pinned-host y -> H2D -> compress ->
            D2H of the packed strings + offsets, every step.
`--impl reference` = the reference's own CPU range coder (oracle/_ref: cc/lib/range_coder.cc compiled in
            place, driven by the restated op loops with all host threads), same symbols and tables.

Weak scaling: every rank codes its own 256-stream batch; rank 0 builds the tables and broadcasts them
(NCCL); there is no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CFG = dict(batch=256, hw=16, channels=128, precision=12, tail_mass=2**-8, n_rot=6)


def _peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      p = json.load(f)
    return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
  return 6650.0, "fallback (B200_PROFILING.md)"


def synth_latents(rank, n_rot, device=None):
  """y[b,h,w,c] ~ Laplace(0, s_c), s_c log-spaced 0.3..8 over channels."""
  import torch
  C, B, HW = CFG["channels"], CFG["batch"], CFG["hw"]
  g = torch.Generator().manual_seed(2 + 1000 * rank)
  scales = torch.exp(torch.linspace(np.log(0.3), np.log(8.0), C))
  out = []
  for _ in range(n_rot):
    u = torch.rand(B, HW, HW, C, generator=g) - 0.5
    y = -scales * torch.sign(u) * torch.log1p(-2 * u.abs())
    out.append(y.contiguous())
  return scales, out


def build_model(scales, device):
  """ContinuousBatchedEntropyModel over per-channel NoisyLaplace priors (exercises the device table
  builder, tfcb_build_lookup)."""
  import torch
  import compression_b200 as tfc
  prior = tfc.NoisyLaplace(loc=torch.zeros_like(scales), scale=scales)
  return tfc.ContinuousBatchedEntropyModel(prior, coding_rank=3, compression=True,
                                           tail_mass=CFG["tail_mass"],
                                           range_coder_precision=CFG["precision"]).to(device)


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
  Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index, uuid=None):
    self.index, self.uuid, self.rows, self._stop = index, uuid, [], threading.Event()
    self._t = threading.Thread(target=self._run, daemon=True)

  def _run_nvml(self):
    """In-process NVML sampling (no nvidia-smi process per sample: those perturb the timed region)."""
    import pynvml
    pynvml.nvmlInit()
    h = None
    if self.uuid:
      for u in (self.uuid, "GPU-" + self.uuid):
        try:
          h = pynvml.nvmlDeviceGetHandleByUUID(u if isinstance(u, bytes) else u.encode())
          break
        except Exception:  # pylint:disable=broad-except
          h = None
    if h is None:
      h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
    mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
    get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
    bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
    while not self._stop.is_set():
      sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
      r = int(get_reasons(h))
      self.rows.append([str(sm), str(mx)] + [("Active" if r & bits[n] else "Not Active")
                                              for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
      self._stop.wait(0.05)

  def _run(self):
    try:
      self._run_nvml()
      return
    except Exception:  # pylint:disable=broad-except
      pass
    while not self._stop.is_set():
      try:
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        self.rows.append([c.strip() for c in out.strip().split(",")])
      except Exception:  # pylint:disable=broad-except
        pass
      self._stop.wait(0.5)

  def __enter__(self):
    self._t.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    self._t.join(timeout=6)

  def summary(self):
    sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
    mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
    return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(sm)}


def cpu_reference(value, lookup, threads, repeats=3):
  """Times the reference CPU range coder (EntropyEncodeChannel + Finalize) on int32 symbols."""
  import oracle
  O = oracle.best()
  S, N = value.shape
  best = None
  for _ in range(repeats):
    enc = O.encoder(lookup, S)
    t0 = time.perf_counter()
    enc.encode(value, None, threads)
    enc.finalize()
    dt = time.perf_counter() - t0
    enc.close()
    best = dt if best is None else min(best, dt)
  return S * N / best / 1e6, O.kind


def symbols_of(model, y):
  """Host int32 symbols exactly as ContinuousBatchedEntropyModel.compress derives them."""
  import torch
  q = model.quantization_offset
  b = y if q is None else y - q.cpu()
  sym = torch.round(b).to(torch.int32) - model.cdf_offset.cpu()
  return sym.reshape(y.shape[0], -1).numpy()


def run_reference(args):
  """--impl reference: rank 0 only; every step codes the full cfg2 batch on the host cores."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import torch
  import oracle
  scales, ys = synth_latents(0, 1)
  cores = os.cpu_count() or 1
  # tables: build on the GPU when there is one (same tables as the main arm), else a Laplace stand-in
  if torch.cuda.is_available():
    model = build_model(scales, "cuda")
    lookup = model.cdf.cpu().numpy()
    value = symbols_of(model, ys[0])
  else:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    cdfs = [util.laplace_cdf(2 * int(6 * s + 3) + 1, CFG["precision"], float(s)) for s in scales]
    lookup = util.make_lookup_1d(cdfs, [CFG["precision"]] * len(cdfs), [True] * len(cdfs))
    off = np.asarray([(len(c) - 1) // 2 for c in cdfs], np.int32)
    value = (torch.round(ys[0]).to(torch.int32).reshape(CFG["batch"], -1).numpy() +
             np.tile(off, CFG["hw"] * CFG["hw"])).astype(np.int32)
  O = oracle.best()
  S, N = value.shape
  for _ in range(args.warmup):
    e = O.encoder(lookup, S); e.encode(value, None, cores); e.finalize(); e.close()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    e = O.encoder(lookup, S); e.encode(value, None, cores); e.finalize(); e.close()
  dt = time.perf_counter() - t0
  val = S * N * args.steps / dt / 1e6
  sample = f"full cfg2 batch ({S} streams x {N} int32 symbols) per step; EntropyEncodeChannel+Finalize only"
  print(json.dumps({
      "impl": "reference", "metric": "range-code throughput (bls2017 compress path, cfg2)", "value": val,
      "unit": "Msymbols/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "u32", "data": "synthetic",
      "config": {"workload": "cfg2: bls2017 compress, y[256,16,16,128], 256 streams x 32768 symbols, 128 tables"},
      "cpu_baseline": {"value": val, "unit": "Msymbols/s", "cores": cores, "kind": O.kind, "sample": sample},
      "e2e": {"value": val, "unit": "Msymbols/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--no-extras", action="store_true", help="skip the decode / GDN / CPU-baseline side measurements")
  args = ap.parse_args()
  if args.impl == "reference":
    return run_reference(args)

  import torch
  import torch.distributed as dist
  import compression_b200 as tfc
  from compression_b200 import _lib, functional, gen_ops

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

  n_rot = CFG["n_rot"]
  scales, ys_host = synth_latents(rank, n_rot)
  # rank 0 builds the tables; everyone else receives them (the only collective on the path)
  from compression_b200 import sharding
  if rank == 0:
    model = build_model(scales, dev)
  else:
    model = tfc.ContinuousBatchedEntropyModel(prior_shape=(CFG["channels"],), coding_rank=3, compression=True,
                                              cdf_shapes=(1, 1), quantization_offset=True).to(dev)
  sharding.broadcast_tables(model, src=0, device=dev)

  ys = [y.to(dev) for y in ys_host]          # > L2: 6 x 33.5 MB rotate through the timed steps
  ys_pinned = [y.pin_memory() for y in ys_host]
  S = CFG["batch"]
  N = ys[0].numel() // S
  sym_per_step = S * N

  def step(i):
    return model.compress(ys[i % n_rot])

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()

  for i in range(max(args.warmup, 3)):
    strings = step(i)
  barrier()
  bits_per_symbol = 8.0 * strings.nbytes() / sym_per_step

  # ---- the timed region: exactly K steps, CUDA events, max over ranks ----
  launches0 = _lib.launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  # clocks are sampled by rank 0 only (its own GPU, in-process NVML): one sampler per rank disturbed the others
  clocks = ClockSampler(local, str(torch.cuda.get_device_properties(dev).uuid)) if rank == 0 else None
  if clocks:
    clocks.__enter__()
  barrier()
  ev0.record()
  for i in range(args.steps):
    strings = step(i)
  ev1.record()
  barrier()
  if clocks:
    clocks.__exit__()
  elapsed_ms = ev0.elapsed_time(ev1)
  launches = _lib.launch_count() - launches0
  t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  elapsed_ms = float(t.item())
  value = world * sym_per_step * args.steps / (elapsed_ms * 1e-3) / 1e6

  # ---- e2e: pinned host y -> H2D -> compress -> D2H(bytes, offsets), same K steps ----
  # Two staging buffers and a copy stream: the H2D copy of batch i+1 runs while batch i is encoded (compress()
  # blocks the host at its finalize, so the next copy has to be queued before it).  Every step's H2D and D2H
  # happen inside the timed region; the result (bytes + offsets) lands in pinned host memory.
  out_cap = 2 * strings.nbytes() + 4096
  host_bytes = [torch.empty(out_cap, dtype=torch.uint8).pin_memory() for _ in range(2)]
  host_offs = [torch.empty(S + 1, dtype=torch.int64).pin_memory() for _ in range(2)]
  stages = [torch.empty_like(ys[0]) for _ in range(2)]
  copy_stream = torch.cuda.Stream(device=dev)
  main_stream = torch.cuda.current_stream()
  ready = [torch.cuda.Event() for _ in range(2)]   # stage[b] holds its batch
  freed = [torch.cuda.Event() for _ in range(2)]   # the encoder is done reading stage[b]

  def queue_h2d(i):
    b = i & 1
    with torch.cuda.stream(copy_stream):
      copy_stream.wait_event(freed[b])
      stages[b].copy_(ys_pinned[i % n_rot], non_blocking=True)
      ready[b].record(copy_stream)

  def e2e_run(k):
    nb = 0
    for b in range(2):
      freed[b].record(main_stream)
    queue_h2d(0)
    for i in range(k):
      b = i & 1
      if i + 1 < k:
        queue_h2d(i + 1)
      main_stream.wait_event(ready[b])
      s = model.compress(stages[b])
      freed[b].record(main_stream)
      nb = s.nbytes()
      host_bytes[b][:nb].copy_(s.bytes_dev[:nb], non_blocking=True)
      host_offs[b].copy_(s.offsets_dev, non_blocking=True)
    main_stream.synchronize()
    return nb

  e2e_run(3)
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  nb = e2e_run(args.steps)
  e1.record()
  barrier()
  e2e_ms = e0.elapsed_time(e1)
  t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  e2e_value = world * sym_per_step * args.steps / (float(t.item()) * 1e-3) / 1e6

  result = {
      "metric": "range-code throughput (bls2017 compress path, cfg2)",
      "value": value, "unit": "Msymbols/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
      "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "u32", "data": "synthetic",
      "config": {
          "workload": "cfg2: bls2017 compress, y[256,16,16,128] fp32 per GPU, 256 streams x 32768 symbols, "
                      "128 NoisyLaplace channel tables, precision 12, overflow on; step = fused quantise + "
                      "range-encode + finalize/pack",
          "bits_per_symbol": round(bits_per_symbol, 4), "streams_per_gpu": S, "symbols_per_stream": N,
          "l2": f"inputs rotate over {n_rot} distinct batches ({n_rot * 33.5:.0f} MB > 126 MB L2)",
          "parallelism": f"batch-shard x{world}, tables broadcast from rank 0",
      },
      "e2e": {"value": e2e_value, "unit": "Msymbols/s", "h2d_bytes_per_step": int(world * ys[0].numel() * 4),
              "d2h_bytes_per_step": int(world * (nb + 8 * (S + 1))),
              "note": "bytes summed over all ranks; H2D of batch i+1 overlaps the encode of batch i (2 staging buffers)"},
      "gpu_launches": int(launches),
  }

  if rank == 0:
    result["clocks"] = clocks.summary()

  if rank == 0 and not args.no_extras:
    peak, peak_src = _peaks()
    # --- dominant kernel of the step: the encode kernel, timed alone with events on the launch stream
    coff = model.cdf_offset.reshape(-1)
    qoff = model.quantization_offset
    lookup = model._lookup_host()
    times = []
    for i in range(6):
      h = gen_ops.create_range_encoder([S], lookup)
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      functional.encode_channel_f32(h, ys[i % n_rot], qoff, coff)
      b.record()
      torch.cuda.synchronize()
      times.append(a.elapsed_time(b))
      h.close()
    enc_ms = float(np.median(times[1:]))
    alg_bytes = sym_per_step * 4 + strings.nbytes()
    achieved = alg_bytes / (enc_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "encode_kernel_traffic.json")
    if os.path.exists(prof):
      with open(prof) as f:
        traffic = json.load(f).get("dram_bytes_per_launch")
    result["roofline"] = {
        "kernel": "encode_kernel (fused quantise + range encode; one CTA per stream: gather, chain and drain warps)", "bound": "hbm",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": peak_src, "kernel_ms": enc_ms, "algorithmic_bytes": alg_bytes,
        "kernel_msym_s": sym_per_step / (enc_ms * 1e-3) / 1e6,
        "note": "latency-bound serial recurrence per stream (256 warps on 148 SMs); HBM fraction is small by construction",
    }
    # --- decode path (create + fused decode/dequantise + finalize), same strings
    strings = model.compress(ys[0])
    for _ in range(2):
      out = model.decompress(strings, (CFG["hw"], CFG["hw"]))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, min(args.steps, 10))
    a.record()
    for _ in range(reps):
      out = model.decompress(strings, (CFG["hw"], CFG["hw"]))
    b.record()
    torch.cuda.synchronize()
    dec_ms = a.elapsed_time(b) / reps
    ok = bool(torch.equal(out, model.quantize(ys[0])))
    result["decode"] = {"value": sym_per_step / (dec_ms * 1e-3) / 1e6, "unit": "Msymbols/s", "ms_per_step": dec_ms,
                        "roundtrip_equals_quantize": ok}
    # --- configs[2]: bmshj2018 hyperprior, batch 128 of 256x256 images: y[128,16,16,192] coded with the indexed
    #     NoisyNormal model (64 scales, per-element index + mean), z[128,4,4,192] with the batched model
    try:
      B3, C3 = 128, 192
      num_scales, smin, smax = 64, .11, 256.
      off3, fac3 = np.log(smin), (np.log(smax) - np.log(smin)) / (num_scales - 1.)
      scale_fn = lambda i: torch.exp(off3 + fac3 * i)
      em_y = tfc.LocationScaleIndexedEntropyModel(tfc.NoisyNormal, num_scales, scale_fn, coding_rank=3, compression=True)
      g3 = torch.Generator(device=dev).manual_seed(7)
      idx3 = torch.rand(B3, 16, 16, C3, device=dev, generator=g3) * 40.0
      sig3 = scale_fn(torch.clamp(idx3, 0, num_scales - 1).to(torch.int32).float())
      loc3 = torch.randn(B3, 16, 16, C3, device=dev, generator=g3)
      y3 = loc3 + sig3 * torch.randn(B3, 16, 16, C3, device=dev, generator=g3)
      zs = torch.exp(torch.linspace(np.log(0.5), np.log(6.0), C3))
      em_z = tfc.ContinuousBatchedEntropyModel(tfc.NoisyLaplace(loc=torch.zeros_like(zs), scale=zs), coding_rank=3,
                                               compression=True).to(dev)
      z3 = (torch.randn(B3, 4, 4, C3, device=dev, generator=g3) * zs.to(dev))
      def enc3():
        return em_z.compress(z3), em_y.compress(y3, idx3, loc=loc3)
      sz, sy = enc3()
      def dec3():
        return em_z.decompress(sz, (4, 4)), em_y.decompress(sy, idx3, loc=loc3)
      res3 = {}
      for name, fn in (("encode", enc3), ("decode", dec3)):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
          out3 = fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        nsym = y3.numel() + z3.numel()
        res3[name] = {"ms_per_step": ms, "value": nsym / (ms * 1e-3) / 1e6, "unit": "Msymbols/s"}
      zhat, yhat = out3
      res3["roundtrip_equals_quantize"] = bool(torch.equal(yhat, em_y.quantize(y3, loc3)) and torch.equal(zhat, em_z.quantize(z3)))
      res3["bits_per_symbol_y"] = 8.0 * sy.nbytes() / y3.numel()
      res3["workload"] = "bmshj2018 two-level: y[128,16,16,192] indexed NoisyNormal (64 scales, loc) + z[128,4,4,192] batched NoisyLaplace; 128 streams each"
      result["cfg3_bmshj2018"] = res3
    except Exception as e:  # pylint:disable=broad-except
      result["cfg3_bmshj2018"] = {"error": repr(e)}
    # --- GDN at the two analysis-transform shapes of cfg2 (forward) and backward at the first
    gdn = {}
    gamma = (0.1 * torch.eye(128) + (0.02 * torch.randn(128, 128)).abs()).to(dev)
    beta = (1 + 0.5 * torch.rand(128)).to(dev)
    for name, npix in (("gdn_0 [256,64,64,128]", 256 * 64 * 64), ("gdn_1 [256,32,32,128]", 256 * 32 * 32)):
      x = torch.randn(npix, 128, device=dev)
      for _ in range(2):
        functional.gdn_forward(x, gamma, beta)
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(5):
        functional.gdn_forward(x, gamma, beta)
      b.record()
      torch.cuda.synchronize()
      ms = a.elapsed_time(b) / 5
      gbs = 8.0 * npix * 128 / (ms * 1e-3) / 1e9
      gdn[name] = {"fwd_ms": ms, "fwd_GBps": gbs, "fwd_frac_of_hbm_peak": gbs / peak}
      if npix == 256 * 64 * 64:
        dy = torch.randn_like(x)
        functional.gdn_backward(x, gamma, beta, dy)
        a.record()
        for _ in range(3):
          functional.gdn_backward(x, gamma, beta, dy)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        gbs = 12.0 * npix * 128 / (ms * 1e-3) / 1e9
        gdn[name].update({"bwd_ms": ms, "bwd_GBps": gbs, "bwd_frac_of_hbm_peak": gbs / peak})
      del x
    # --- configs[3]: GDN microbench, 192 channels, 64x64 tiles (batch 4096 if memory allows, else 1024)
    try:
      free_b, _ = torch.cuda.mem_get_info(dev)
      batch4 = 4096 if free_b > 90e9 else 1024
      npix = batch4 * 64 * 64
      gamma192 = (0.1 * torch.eye(192) + (0.02 * torch.randn(192, 192)).abs()).to(dev)
      beta192 = (1 + 0.5 * torch.rand(192)).to(dev)
      x = torch.randn(npix, 192, device=dev) * (0.05 + 3.95 * torch.rand(192, device=dev))  # SURVEY 8(d) cfg4 recipe
      functional.gdn_forward(x, gamma192, beta192)
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(3):
        functional.gdn_forward(x, gamma192, beta192)
      b.record()
      torch.cuda.synchronize()
      ms = a.elapsed_time(b) / 3
      gbs = 8.0 * npix * 192 / (ms * 1e-3) / 1e9
      entry = {"fwd_ms": ms, "fwd_GBps": gbs, "fwd_frac_of_hbm_peak": gbs / peak, "fwd_kernel": "tcgen05 (gdn_tc_fwd3_kernel<192>)"}
      dy = torch.randn_like(x)
      functional.gdn_backward(x, gamma192, beta192, dy)
      a.record()
      functional.gdn_backward(x, gamma192, beta192, dy)
      b.record()
      torch.cuda.synchronize()
      ms = a.elapsed_time(b)
      gbs = 12.0 * npix * 192 / (ms * 1e-3) / 1e9
      entry.update({"bwd_ms": ms, "bwd_GBps": gbs, "bwd_frac_of_hbm_peak": gbs / peak, "bwd_kernel": "tcgen05, two kernels (gdn_tc_bwd_dx_kernel<192> + gdn_tc_bwd_dgamma_kernel<192>)"})
      gdn[f"cfg4 [{batch4},64,64,192]"] = entry
      del x, dy
    except Exception as e:  # pylint:disable=broad-except
      gdn["cfg4 [4096,64,64,192]"] = {"error": repr(e)}
    result["gdn"] = gdn
    # --- CPU baseline: the reference's own range coder on this box's host cores (bounded: one batch)
    try:
      value_host = symbols_of(model, ys_host[0])
      cores = os.cpu_count() or 1
      all_cores, kind = cpu_reference(value_host, lookup, cores)
      one_core, _ = cpu_reference(value_host[:16], lookup, 1, repeats=2)
      result["cpu_baseline"] = {
          "value": all_cores, "unit": "Msymbols/s", "cores": cores, "kind": kind,
          "sample": f"one full cfg2 batch ({S} streams x {N} int32 symbols), EntropyEncodeChannel+Finalize, "
                    f"best of 3, streams sharded over {cores} threads",
          "single_core_value": one_core, "single_core_sample": "16 streams x 32768 symbols, 1 thread",
      }
    except Exception as e:  # pylint:disable=broad-except
      result["cpu_baseline"] = {"error": repr(e)}

  if rank == 0:
    print(json.dumps(result))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
