#!/usr/bin/env python
"""bench.py -- pixel-iterations/s of the escape-time path on B200.

One "step" is one complete render of the workload view (reset queue + the
persistent escape-time kernel).  At N=1 the workload is BASELINE.json
configs[1]: the full Mandelbrot set, 1920x1080, MDZ's hardware-precision mode
(x87 long double == 64-bit significand, SURVEY finding 1), depth 10000.  With N
ranks (torchrun, one process per GPU) the same image is split into interleaved
line bands, one plan per rank, no data-path collective ("strong" scaling).

Prints ONE JSON line (rank 0).  `--impl reference` instead times the unmodified
reference's own pthread pool (oracle/_ref/libmdzref.so) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "pixel_iterations_per_sec"
UNIT = "pixel-iterations/s"


def macs_per_iteration(n_limbs):
    """one general product N^2 + two squarings N(N+1)/2 each."""
    return 2 * n_limbs * n_limbs + n_limbs


def workload_cfg4(mode="mpfr"):
    """BASELINE configs[3] / the north star's target: a 1e-120 wide view at 512 bits, 3840x2160,
    depth 100000 (tests/views.py config4; every pixel escapes after ~12 800 iterations), in MPFR
    mode (the north star's "512-bit MPFR-equivalent") or GMP mpf mode (as configs[3] names it).
    One image split across the ranks by interleaved bands: strong scaling."""
    from views import config4
    return config4(3840, 2160, 100000, mode=mode, precision=512)


def workload_view(world=1, scaling="weak"):
    """BASELINE configs[1] at N=1.  With N ranks the path shards by line bands
    (no collective), so by default the benchmark is weak-scaled: the same view at
    N times the pixels (1920x1080 per GPU: 2716x1528 on 2, 3840x2160 on 4,
    5432x3056 on 8), each rank rendering bands r, r+N, ...  `--scaling strong`
    keeps the 1920x1080 image and splits it instead."""
    from views import config2
    if world == 1 or scaling == "strong":
        return config2(1920, 1080, 10000)
    f = world ** 0.5
    w = int(round(1920 * f / 8.0)) * 8
    h = int(round(1080 * f / 8.0)) * 8
    return config2(w, h, 10000)


def iterations_of(raw, depth):
    import numpy as np
    return int(np.where(raw > 0, raw, depth).astype(np.int64).sum())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, emit):
    """Reference arm: the unmodified reference pool on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import refpath
    from views import config2
    lib = refpath.load()
    cores = os.cpu_count() or 1
    # bounded sample: the same view at half resolution each way (same set, 1/4 of the pixels)
    view = config2(960, 540, 10000)
    kind = "reference"
    if lib is None:
        import portpath
        kind = "port"
    times, iters = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if lib is not None:
            raw, _ = refpath.ref_render(lib, view, cores)
        else:
            raw = portpath.port_render(view, cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            iters = iterations_of(raw, view.depth)
    total = sum(times)
    value = iters * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "x87 long double (64-bit significand)",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: full M-set, long double, depth 10000 "
                               "(sample rendered at 960x540)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "same view at 960x540 (1/4 of the pixels), %d pixel-iterations per step, "
                                   "reference pthread pool with -t %d" % (iters, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def side_precisions(torch, device):
    """Kernel-only it/s at the other precisions of the metric, on bounded views."""
    import mdz_b200
    from views import make_view, SEAHORSE, deep_embedded_julia
    out = {}
    cases = [
        ("mpfr128", make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", 1920, 1080, precision=128, depth=10000)),
        ("mpfr320", deep_embedded_julia(960, 540)),
        ("mpfr512", make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", 960, 540, precision=512, depth=10000)),
        ("gmp512", make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", 960, 540, mode="gmp", precision=512, depth=10000)),
    ]
    peak = mdz_b200.imad_peak(device, 100)
    peak32 = mdz_b200.imad_peak(device, 100, wide=False)
    for name, view in cases:
        plan = mdz_b200.Plan(view, device)
        st = torch.cuda.current_stream().cuda_stream
        plan.launch(st); plan.wait()                    # warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.launch(st); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        iters = iterations_of(plan.fetch(), view.depth)
        ki = plan.kernel_info()
        rate = iters / (ms * 1e-3)
        # GMP mode multiplies the top P 64-bit limbs of each operand: N = 2P words
        n_mac = ki["limbs"] - 2 if view.mode == 2 else ki["limbs"]
        out[name] = {"value": rate, "unit": UNIT, "ms": ms, "pixel_iterations": iters,
                     "view": "%dx%d" % (view.real_width, view.real_height),
                     "limbs": ki["limbs"], "mac_limbs": n_mac, "macs_per_iteration": macs_per_iteration(n_mac),
                     "regs": ki["regs_per_thread"], "spill_bytes": ki["local_bytes"],
                     "blocks_per_sm": ki["blocks_per_sm"],
                     "imad_frac": rate * macs_per_iteration(n_mac) / peak,
                     "frac_of_imad32_issue": rate * macs_per_iteration(n_mac) / peak32}
        plan.close()
    return out


def main():
    # rank 0 must print exactly ONE line on stdout; NCCL / torch may write banners to fd 1
    # ("NCCL version ..."), so park the real stdout and send everything else to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-side", action="store_true", help="skip the per-precision side measurements")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4", "cfg4gmp"],
                    help="cfg2 (default, the bench contract's workload): BASELINE configs[1]; cfg4: the north star's "
                         "target view, 3840x2160 at 512-bit MPFR, depth 100000, one image split over the ranks")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = same view at N x the pixels (default), strong = split the 1920x1080 image")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args, emit)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import mdz_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg4 = args.workload in ("cfg4", "cfg4gmp")
    gmp4 = args.workload == "cfg4gmp"
    if cfg4:
        args.scaling = "strong"
        args.no_side = True
    view = workload_cfg4("gmp" if gmp4 else "mpfr") if cfg4 else workload_view(world, args.scaling)
    plan = mdz_b200.Plan(view, local, band_first=rank, band_stride=world)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    for _ in range(args.warmup):
        flush.zero_()
        plan.launch(stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launched0 = plan.kernels_launched()
    e0.record()
    for i in range(args.steps):
        flush.zero_()
        kev[i][0].record()
        plan.launch(stream)
        kev[i][1].record()
    e1.record()
    barrier()
    my_launches = plan.kernels_launched() - launched0      # 1 per render, 3 with tail compaction
    ms_total = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    raw = plan.fetch()
    my_lines = raw[[l for b in range(rank, view.user_height, world)
                    for l in range(b * view.aa_factor, (b + 1) * view.aa_factor)]]
    my_iters = iterations_of(my_lines, view.depth)
    t = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device="cuda")
    it = torch.tensor([my_iters, my_launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
    ms_total, kernel_ms = float(t[0]), float(t[1])
    total_iters = int(it[0])
    total_launches = int(it[1])
    value = total_iters * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the public call: host view in, host raw_data out ----
    xs_bytes = (plan.kernel_info()["limbs"] + 2) * 4 * (view.real_width + plan.local_lines() + 2)
    e2e_steps = max(1, min(args.steps, 2 if cfg4 else 20))
    out = np.empty((view.real_height, view.real_width), dtype=np.int32)   # the caller's raw_data (pageable, as MDZ's malloc)
    out.fill(-1)
    for _ in range(1 if cfg4 else 2):                                           # untimed: pool warm, pages touched
        p2 = mdz_b200.Plan(view, local, band_first=rank, band_stride=world); p2.run(out, stream); p2.close()
    barrier()
    e2e_ms = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        p2 = mdz_b200.Plan(view, local, band_first=rank, band_stride=world)   # prologue + H2D of the tables
        p2.run(out, stream)                                                      # kernel; bands D2H as they complete
        p2.close()
        e2e_ms.append(round((time.perf_counter() - ta) * 1e3, 3))
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_iters * e2e_steps / float(te[0])
    assert np.array_equal(out[rank * view.aa_factor], raw[rank * view.aa_factor])

    if rank == 0:
        ki = plan.kernel_info()
        # per-launch figures of this kernel from the committed ncu capture (profiles/): DRAM
        # traffic and which pipe binds.  Static facts of the build, not measured in this run.
        ncu, ncu1 = {}, {}
        try:
            with open(os.path.join(ROOT, "profiles", "r1_ncu_summary.json")) as f:
                both = json.load(f)
            ncu = both.get("ld64_cfg2_phase0", {})       # phase 0: ~80 % of a render (profiles/r1_launches.csv)
            ncu1 = both.get("ld64_cfg2_phase1", {})      # phase 1: the parked pixels (tail compaction)
        except (OSError, ValueError):
            pass

        def ncu_num(key, src=None):
            try:
                v, unit = (ncu if src is None else src)[key].split()[:2]
                return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            except (KeyError, ValueError, IndexError):
                return None
        dram = None
        parts = [ncu_num(k, src) for src in (ncu, ncu1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
        if all(x is not None for x in parts):
            dram = sum(parts)
        peak = mdz_b200.imad_peak(local, 200)                  # IMAD.WIDE.U32.X chains: 32x32->64 MAC/s
        peak32 = mdz_b200.imad_peak(local, 200, wide=False)    # 32-bit IMAD issue rate
        macs = macs_per_iteration(ki["limbs"] - 2 if view.mode == 2 else ki["limbs"])     # GMP mode multiplies the top P 64-bit limbs: N = 2P words
        kernel_rate = (total_iters / world) / (kernel_ms * 1e-3)        # this GPU's kernel alone
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": args.scaling if (world > 1 or cfg4) else "weak", "vs_baseline": None,
            "dtype": ("20 x u32 words (9 + 1 64-bit limbs) + limb exponent (soft-float == GMP 6.3 mpf at 512 bits, truncating)" if gmp4 else
                      "16 x u32 significand + i32 exponent (soft-float == MPFR at 512 bits, round to nearest even)" if cfg4 else
                      "u64 significand + i32 exponent (soft-float == x87 long double, round to nearest even)"),
            "data": "synthetic",
            "config": {"workload": ("BASELINE configs[3] / north-star target: 1e-120 wide view on M(23,2), 3840x2160, %s 512 bits, "
                                    "depth 100000, one image split over %d GPU(s) (strong scaling)" % ("GMP mpf" if gmp4 else "MPFR", world)) if cfg4 else
                                   "BASELINE configs[1]: full M-set cx=-0.5 cy=0 size=4, %dx%d, "
                                   "long double mode, depth 10000%s" % (
                                       view.real_width, view.real_height,
                                       "" if world == 1 else (" (weak scaling: 1920x1080 pixels per GPU)" if args.scaling == "weak"
                                                               else " (strong scaling: one 1920x1080 image split)")),
                       "pixel_iterations_per_step": total_iters,
                       "partition": "interleaved line bands, one plan per rank, no collective",
                       "l2": "256 MiB buffer rewritten between steps (inputs are KB-sized tables)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                    "h2d_bytes_per_step": xs_bytes, "d2h_bytes_per_step": int(my_lines.nbytes),
                    "ms_each": e2e_ms,
                    "includes": "mdzcuda_plan_create (host prologue with libmpfr/long double, H2D of the tables) + "
                                "mdzcuda_plan_run (kernel, D2H of finished bands into pageable host raw_data) + destroy"},
            "gpu_launches": total_launches,
            "clocks": clocks,
            "roofline": {"bound": "imad", "achieved": kernel_rate * macs / 1e12, "peak": peak / 1e12,
                         "unit": "T 32x32->64 MAC/s", "frac": kernel_rate * macs / peak,
                         "traffic": dram,
                         "traffic_note": "dram__bytes_read+write of one render = both launches of the escape kernel (ncu --set full, "
                                         "profiles/r1_ncu_summary.json: ld64_cfg2_phase0 + ld64_cfg2_phase1): 0.18 MB in phase 0, 9.2 MB in "
                                         "phase 1 reading the parked states back; the 8.3 MB of results stay in L2 until after the kernel",
                         "binding_pipe": {"pipe": "alu", "busy_pct": ncu_num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                                          "fma_pct": ncu_num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                                          "fmaheavy_cycles_pct": ncu_num("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                                          "fp64_pct": ncu_num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                                          "issue_active_pct": ncu_num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                          "source": "ncu capture of this kernel on this workload (profiles/r1_ncu_summary.json: ld64_cfg2_phase0, "
                                                    "the launch that is ~80 % of a render)",
                                          "why": "at 64 bits an iteration is 12 IMAD.WIDE against ~170 shift/compare/select/add "
                                                 "instructions of alignment, normalisation and rounding: the ALU pipe binds, not the multiplier"},
                         "peak_imad32": peak32 / 1e12,
                         "frac_of_imad32_issue": kernel_rate * macs / peak32,
                         "note": "integer-pipe roofline (SURVEY 8d), not hbm/tensor. achieved = it/s x W(N)=2N^2+N "
                                 "full-schoolbook MACs (N=%d limbs -> %d; the kernel forms only the high ~59%% of each "
                                 "product, DESIGN.md 2.1); peak = IMAD.WIDE.U32.X carry-chain microbenchmark measured in "
                                 "this run (MEASURED_PEAKS.json has no integer peak); peak_imad32 = 32-bit IMAD issue rate, "
                                 "twice that; kernel avg %.3f ms per render by CUDA events (%d launch(es) per render: with tail "
                                 "compaction the escape kernel runs twice around a one-block ordering pass); HBM traffic is 4 B per pixel out"
                                 % (ki["limbs"], macs, kernel_ms, total_launches // max(1, args.steps * world))},
            "kernel": ki,
        }
        if cfg4:
            # the committed ncu figures are the long double kernel's: for this workload say what binds from
            # the 512-bit capture instead (profiles/r1_ncu_summary.json: mpfr512_seahorse_960x540)
            line["roofline"]["traffic"] = None
            line["roofline"].pop("traffic_note", None)
            line["roofline"]["binding_pipe"] = {"pipe": "fmaheavy + alu", "source": "no ncu capture of the GMP kernel this round",
                                                "why": "escape_gmpf_kernel<20>: 666 MACs per iteration in full-schoolbook terms"} if gmp4 else {
                "pipe": "fmaheavy + alu (issue-limited)", "source": "profiles/r1_ncu_summary.json: mpfr512_seahorse_960x540",
                "why": "escape_mpfr_kernel<16>: 1281 issue slots per iteration of which 311 IMAD.WIDE; fmaheavy 57 %, ALU 54 %, "
                       "issue slots 51 % busy at 3 warps per scheduler (168 registers)"}
        # CPU baseline: the unmodified reference pool on this box's cores, bounded sample
        try:
            import refpath
            from views import config2
            lib = refpath.load()
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            if cfg4:
                lines = [(2 * k + 1) * view.real_height // (2 * cores) for k in range(cores)]    # one line per thread
                rraw = refpath.ref_render_lines(lib, view, lines, cores)
                sample_txt = "%d evenly spaced lines of the same 3840x2160 view, the reference's own line driver" % len(lines)
                sdepth = view.depth
                same = bool(np.array_equal(rraw, raw[lines])) if world == 1 else None
            else:
                sample = config2(960, 540, 10000)
                rraw, _ = refpath.ref_render(lib, sample, cores)
                sample_txt = "same view at 960x540, unmodified reference pool"
                sdepth = sample.depth
                same = None
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": iterations_of(rraw, sdepth) / dt, "unit": UNIT,
                                    "cores": cores, "kind": "reference",
                                    "sample": "%s, -t %d, %.2f s" % (sample_txt, cores, dt)}
            if same is not None:
                line["cpu_baseline"]["identical_to_gpu_on_sample"] = same
        except Exception as ex:   # the reference .so is optional on the box
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
        if not args.no_side and world == 1:
            line["per_precision"] = side_precisions(torch, local)
            # NOT the metric: the same render with the exact periodicity check on (include/mdzcuda.h),
            # i.e. what a user of the drop-in gets.  raw_data is identical; iterations performed are not
            # the reference's count any more, so no it/s figure is derived from it.
            ms = []
            for _ in range(4):
                ta = time.perf_counter()
                p3 = mdz_b200.Plan(view, local)
                p3.set_cycle_detection(True)
                got = p3.run(stream=stream)
                p3.close()
                ms.append((time.perf_counter() - ta) * 1e3)
            line["cycle_detection"] = {"e2e_ms": min(ms[1:]), "e2e_ms_full_iteration": min(e2e_ms),
                                       "identical_raw_data": bool(np.array_equal(got, raw)),
                                       "note": "opt-in exact periodicity check; value / e2e above are measured with it off"}
        emit(line)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
