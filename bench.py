#!/usr/bin/env python
"""bench.py — images/sec of the TF2 quantised-convolution hot path on B200 (BASELINE.json metric).

Workload (default, N=1): ResNet50 INT4-weight / INT8-feature, 224x224, batch 256 per GPU (BASELINE
configs[1] with --variant shift, configs[2] with --variant mma; default `auto` = per-layer choice).
`--net vgg16` (configs[3]: batch 1024 over 8 GPUs = 128 per GPU) and `--net googlenet` (configs[4]: batch
512 over 8 = 64 per GPU), `--net squeezenet` (configs[0]'s network) run the same measurement on those
networks.  Synthetic INQ-style weights in the reference's param.bin format, the shipped per-channel Q table
(a synthetic one where the reference ships none), synthetic int8 images.  One "step" = one batch through the
whole network.

  python bench.py --gpus N --steps K --warmup W [--net NET] [--variant auto|shift|mma] [--impl reference]
For N>1 launch with torch.distributed.run (one rank per GPU); images shard across ranks (weak scaling, no
data-path collective; one NCCL broadcast of the weight blob at init).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM), `e2e` the same
through the host-buffer C-ABI call (pinned host int8 images in, logits out, H2D/D2H inside the timed region).
`roofline` describes the dominant kernel family; `cpu_baseline` is the CPU oracle on a bounded sample.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tf2_b200 import capi, formats, nets, synth  # noqa: E402

# per network: batch per GPU of the BASELINE config, pretty name, whether the input goes through the 7x7 -> 3x3 stem
# transform (raw 3x224x224 in, tf2b_run_raw224) or is tensor 0 itself
NETS = {
    "resnet50": dict(batch=256, title="ResNet50", raw224=True),
    "googlenet": dict(batch=64, title="GoogLeNet", raw224=True),
    "vgg16": dict(batch=128, title="VGG16", raw224=False),
    "squeezenet": dict(batch=256, title="SqueezeNet", raw224=False),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def build_model(net_name="resnet50", seed=3):
    net = nets.load(net_name)
    qpath = os.path.join(ROOT, "tests", "golden", f"{net_name}_Q")
    if os.path.exists(qpath):
        q = formats.parse_q_file(net, qpath)
    else:                                   # no table shipped by the reference for this network
        q = formats.parse_q_text(net, synth.synth_q_text(net, seed=seed))
    blob = synth.synth_float_blob(net, seed=seed, q=q)
    model = formats.load_float_blob(net, blob, q)
    return net, q, model


def algorithmic_bytes_per_image(net):
    """Layer-by-layer int8 activation traffic (inputs + outputs + residual reads), SURVEY.md 8d."""
    tot = 0
    for ld in net.layers:
        ti, to = net.tensors[ld.in_tensor], net.tensors[ld.out_tensor]
        cin = ti.C if ld.ipool else ld.C
        tot += cin * ti.H * ti.W
        tot += ld.N * to.H * to.W
        if ld.add_tensor >= 0:
            tot += ld.N * ld.PH * ld.PW
    return tot


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 7:
                for i, nme in enumerate(names):
                    if r[3 + i].lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def input_images(net, cfg, q, n, seed):
    """int8 tensor-0 images for the CPU oracle (raw images through feature_trans where the net has the stem)."""
    imgs = synth.synth_images(n, seed=seed)
    if cfg["raw224"]:
        return formats.prepare_input(net, imgs, q)[1]
    return formats.quantize_input(imgs, int(q[0, 0]))


def cpu_reference_run(net, model, t0, n_threads):
    from oracle import oracle as O
    t = time.perf_counter()
    O.run_network(net, model, t0, n_threads=n_threads)
    return time.perf_counter() - t


def cpu_baseline(net, cfg, q, model, batch, target_s=12.0):
    """CPU oracle (a C restatement = 'port') on a bounded sample of the same workload."""
    cores = len(os.sched_getaffinity(0))
    t0 = input_images(net, cfg, q, max(2, min(cores, 8)), 21)
    dt = cpu_reference_run(net, model, t0, cores)           # warm-up + calibration
    per_img = dt / t0.shape[0]
    n = int(max(2, min(batch, target_s / max(per_img, 1e-6))))
    t0 = input_images(net, cfg, q, n, 22)
    dt = cpu_reference_run(net, model, t0, cores)
    return {"value": n / dt, "unit": "images/sec", "cores": cores, "kind": "port",
            "sample": f"{n} of the {batch} images of one batch, whole {cfg['title']}, oracle/tf2_oracle.c with OpenMP"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--net", default="resnet50", choices=sorted(NETS))
    ap.add_argument("--variant", default="auto", choices=["auto", "shift", "mma"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU and step (default: the BASELINE config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--executor", type=int, default=1, help="tf2b_set_graph: 1 graph + lanes (default), 2 graph, 3 lanes, 0 plain")
    ap.add_argument("--stem-chunk", type=int, default=0, help="chunked L2-resident stem (default off: measured slower)")
    ap.add_argument("--weights", default="planes", choices=["planes", "packed4"], help="tensor-core weight staging")
    ap.add_argument("--layers-out", default=None, help="write per-layer device times (JSON) to this file")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = NETS[args.net]
    B = args.batch if args.batch > 0 else cfg["batch"]
    metric = f"images/sec {cfg['title']}-INT4 224x224"
    workload = f"{cfg['title']} INT4/INT8 batch={B} per GPU, 224x224, variant={args.variant}"
    # identical in both arms (the driver compares them): what is measured, not how
    config = {"workload": workload, "global_batch": B * world, "parallelism": f"dp{world}"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        net, q, model = build_model(args.net)
        cores = len(os.sched_getaffinity(0))
        per_step = max(2, min(cores, 16, B))   # bounded sample: images per step
        t0 = input_images(net, cfg, q, per_step, 5)
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_reference_run(net, model, t0, cores)
        t = time.perf_counter()
        for _ in range(args.steps):
            cpu_reference_run(net, model, t0, cores)
        dt = time.perf_counter() - t
        val = per_step * args.steps / dt
        line = {"metric": metric, "value": val, "unit": "images/sec", "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8xint4->int32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "images/sec", "cores": cores, "kind": "port",
                                 "sample": f"{per_step} images of the batch per step, whole {cfg['title']}, oracle/tf2_oracle.c (C "
                                           f"restatement of the reference device kernels, pinned against them compiled and executed; "
                                           f"the reference's own device program as C is a cycle-level emulation, ~0.03 images/s, "
                                           f"DESIGN.md 5)"},
                "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    from tf2_b200.network import NetWork, Runner
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: tf2_b200 has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    variant = {"auto": capi.VARIANT_AUTO, "shift": capi.VARIANT_SHIFT, "mma": capi.VARIANT_MMA}[args.variant]

    net = nets.load(args.net)
    nw = NetWork(net, device=local_rank)
    nw.set_stem_chunk(bool(args.stem_chunk))
    nw.set_weight_staging(capi.WEIGHTS_PACKED4 if args.weights == "packed4" else capi.WEIGHTS_PLANES)
    q = model = None
    if rank == 0:
        net, q, model = build_model(args.net)
    if world > 1:
        # single NCCL broadcast of the packed weight blob at init
        from tf2_b200.dist import init_network_distributed
        init_network_distributed(nw, dist, dev, model=model, q=q, max_images=B, variant=variant)
    else:
        nw.InitFromCodes(model, q, max_images=B, variant=variant)
    nw.set_graph(args.executor)
    runner = Runner(nw)
    raw224 = cfg["raw224"]
    t0d = net.tensors[0]
    in_shape = (B, 3, 224, 224) if raw224 else (B, t0d.C, t0d.H, t0d.W)
    tres = net.tensors[net.result_tensor()]
    out_shape = (B, tres.C, tres.H, tres.W)

    # synthetic int8 images; 4 distinct batches rotate so inputs are never L2-resident
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    nrot = 4
    host_batches = [torch.randint(-128, 128, in_shape, dtype=torch.int8, generator=g).pin_memory() for _ in range(nrot)]
    dev_batches = [hb.to(dev) for hb in host_batches]
    in_bytes = int(np.prod(in_shape))
    out = torch.empty(out_shape, dtype=torch.int8, device=dev)
    out_host = torch.empty(out_shape, dtype=torch.int8).pin_memory()
    # a stream of our own: the engine's default executor (CUDA graph replay) needs a capturable stream, which the
    # legacy default stream is not; every device-timed launch and both timing events go to this stream
    torch.cuda.synchronize(dev)              # the input batches were copied on the default stream
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        runner.run_device(dev_batches[i % nrot], out=out, raw224=raw224, stream=stream)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        runner.run_device(dev_batches[i % nrot], out=out, raw224=raw224, stream=stream)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = nw.last_launches() * args.steps
    clocks = sampler.stop() if sampler else None

    # end to end through the host-buffer C-ABI call (H2D of the int8 images + D2H of the result inside the timed
    # region, every step).  The public host API is asynchronous like the reference's (EnqueueKernels /
    # WaitForAllKernels): two slots, so the copies of step i+1 / i-1 overlap step i.  Both slots and the
    # synchronous call are warmed up first (all staging memory exists since tf2b_finalize).
    out_hosts = [torch.empty(out_shape, dtype=torch.int8).pin_memory() for _ in range(2)]
    for i in range(max(2, min(args.warmup, 4))):
        runner.submit_host(host_batches[i % nrot], out_hosts[i % 2], i % 2, raw224=raw224)
        runner.wait(i % 2)
    runner.run_host(host_batches[0], out_host, raw224=raw224)
    barrier()
    t = time.perf_counter()
    for i in range(args.steps):
        runner.submit_host(host_batches[i % nrot], out_hosts[i % 2], i % 2, raw224=raw224)
        if i >= 1:
            runner.wait((i - 1) % 2)
    runner.wait((args.steps - 1) % 2)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t
    # the synchronous one-call form, for comparison
    barrier()
    t = time.perf_counter()
    for i in range(args.steps):
        runner.run_host(host_batches[i % nrot], out_host, raw224=raw224)
    torch.cuda.synchronize(dev)
    e2e_sync_s = time.perf_counter() - t

    # per-layer device times of the conv kernels (CUDA events recorded by the library on the launching stream
    # between the layers; they break the programmatic overlap of consecutive layers, so this pass only yields
    # each kernel's SHARE of the step — absolute times come from the un-profiled pass above)
    nw.set_profile(True)
    conv_ms = np.zeros(net.num_layers)
    layer_ms = np.zeros(net.num_layers)
    nprof = 3
    for i in range(nprof):
        runner.run_device(dev_batches[i % nrot], out=out, raw224=raw224, stream=stream)
        c, l = nw.get_profile()
        conv_ms += c
        layer_ms += l
    conv_ms /= nprof
    layer_ms /= nprof
    nw.set_profile(False)
    kernels = nw.layer_kernels()

    def layer_macs(ld):
        if ld.ipool:
            return 0
        if ld.first_layer_7x7:
            return ld.OH * ld.OW * ld.N * 147
        return ld.OH * ld.OW * ld.N * ld.C * ld.k * ld.k

    if args.layers_out and rank == 0:
        modes_l = nw.layer_modes(B)
        rows = []
        for l, ld in enumerate(net.layers):
            rows.append({"layer": l, "kernel": kernels[l], "mode": modes_l[l], "C": ld.C, "N": ld.N, "k": ld.k,
                         "stride": ld.stride, "OH": ld.OH, "conv_ms": float(conv_ms[l]), "layer_ms": float(layer_ms[l]),
                         "tops": (2 * layer_macs(ld) * B / (conv_ms[l] * 1e-3) / 1e12) if conv_ms[l] > 0 else None})
        with open(args.layers_out, "w") as f:
            json.dump(rows, f, indent=0)

    if world > 1:
        tt = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    total_images = B * world * args.steps
    value = total_images / (ms * 1e-3)
    step_ms = ms / args.steps
    macs = net.macs_per_image()
    # dominant kernel family by device time
    fam_time = {}
    for l, kname in enumerate(kernels):
        fam_time[kname] = fam_time.get(kname, 0.0) + float(conv_ms[l])
    dom = max((k for k in fam_time if k != "none"), key=lambda k: fam_time[k])
    dom_layers = [l for l, kname in enumerate(kernels) if kname == dom]
    dom_ops = 2.0 * sum(layer_macs(net.layers[l]) for l in dom_layers) * B
    prof_step_ms = float(layer_ms.sum())
    share = fam_time[dom] / prof_step_ms
    # the kernel family's time inside the timed (un-profiled) steps = its share x the measured step time
    dom_s = share * step_ms * 1e-3
    # INT8 tcgen05 rate = 2x bf16 on sm_100a (tools/micro/umma_rate.cu).  The timed region of the default run is
    # tens of milliseconds at full clocks: the burst figure is the denominator; the sustained one only when the
    # region is long enough to sit in the power-limited regime
    burst = ms < 1000.0
    int8_peak_burst = 2.0 * peaks["bf16_tflops"]
    int8_peak_sust = 2.0 * peaks["bf16_tflops_sustained"]
    int8_peak = int8_peak_burst if burst else int8_peak_sust
    if dom == "shift":
        # CUDA-core kernel: 4-way int8 dot products (IDP.4A) issue on one of the two math pipes, 64 lanes/clk/SM
        sm_clk = (clocks or {}).get("sm_max_mhz") or 1965.0
        int8_peak = int8_peak_burst = int8_peak_sust = 148 * 64 * 4 * 2 * sm_clk * 1e6 / 1e12
    achieved = dom_ops / dom_s / 1e12
    # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this command
    # (profiles/r02_ncu_traffic*.json, written by tools/ncu_traffic.py); null when absent
    traffic = None
    for name in (f"r02_ncu_traffic_{args.net}.json", "r02_ncu_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("kernel") == f"conv_{dom}" and tj.get("batch") == B and tj.get("net", "resnet50") == args.net:
                traffic = tj.get("dram_bytes_per_launch")
                break
    modes = nw.layer_modes(B)
    abytes = algorithmic_bytes_per_image(net)
    roofline = {"bound": "tensor", "kernel": f"conv_{dom}", "achieved": achieved, "peak": int8_peak, "unit": "TFLOP/s",
                "ops_kind": "int8 x int8 -> int32 multiply-accumulate counted as 2 ops (true convolution, conv1 as 7x7x3)",
                "frac": achieved / int8_peak,
                "frac_vs_burst_peak": achieved / int8_peak_burst, "frac_vs_sustained_peak": achieved / int8_peak_sust,
                "peak_source": (f"2 x {peaks['src']} {'burst' if burst else 'sustained'} bf16 cuBLAS TF/s (INT8 MMA issues at "
                                f"twice the bf16 rate); timed region {ms:.0f} ms") if dom != "shift" else
                               "148 SMs x 64 lanes x IDP.4A (4 MAC) x max SM clock",
                "launches_per_step": len(dom_layers),
                "launch_avg_us": 1e6 * dom_s / max(1, len(dom_layers)),
                "launch_avg_us_profiled_pass": 1e3 * fam_time[dom] / max(1, len(dom_layers)),
                "algorithmic_ops_per_launch": dom_ops / max(1, len(dom_layers)),
                "share_of_step": share,
                "step_ms_profiled_pass": prof_step_ms,
                "traffic": traffic,
                "staging_modes": {k: sum(1 for m in modes if k in m) for k in ("flat", "box", "halo", "wstream", "ctapair", "fold", "wres")},
                "hbm_view": {"algorithmic_bytes_per_image": abytes,
                             "achieved_gbs": abytes * B / (step_ms * 1e-3) / 1e9,
                             "peak_gbs": peaks["hbm_gbs"]}}

    line = {"metric": metric, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8xint4->int32", "data": "synthetic",
            "config": config,
            "details": {"net": args.net, "executor": args.executor, "stem_chunk": args.stem_chunk, "weights": args.weights, "kernels": {k: kernels.count(k) for k in sorted(set(kernels))},
                        "l2": f"inputs rotate over {nrot} distinct batches ({nrot * in_bytes / 1e6:.0f} MB); activations per step are GBs",
                        "gmac_per_image": macs / 1e9},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": total_images / e2e_s, "unit": "images/sec",
                    "h2d_bytes_per_step": in_bytes * world, "d2h_bytes_per_step": int(np.prod(out_shape)) * world,
                    "api": ("tf2b_submit_raw224_host" if raw224 else "tf2b_submit_host") + " + tf2b_wait (2 slots, pinned host buffers)",
                    "sync_call_value": B * args.steps / e2e_sync_s},
            "roofline": roofline,
            "frac_of_int8_mma_roofline": (value / world) * 2 * macs / 1e12 / int8_peak,
            "frac_of_hbm_roofline": (value / world) * abytes / 1e9 / peaks["hbm_gbs"]}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(net, cfg, q, model, B)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
