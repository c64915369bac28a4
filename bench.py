#!/usr/bin/env python
"""bench.py -- pixel-iterations/s of the escape-time path on B200.

Default workload, the same at every N: the north star's target -- ONE 3840x2160 image of a
1e-120 wide view at 512-bit MPFR precision, depth 100000, centred next to a minibrot so that
2 % of the pixels run to depth (BASELINE configs[3] geometry, tests/views.py config4m).  One
"step" is one complete render of it.  With N ranks (torchrun, one process per GPU) the image
is split into interleaved line bands, one plan per rank, no data-path collective: strong
scaling.  Untimed, inside the same run: every rank checks sampled lines of its share against
the reference's own line driver, rank 0 renders the image once more through the library's own
multi-device call (mdzcuda_render over all N devices from one process), and at N = 1 the other
precisions of the metric are measured (`per_precision`, BASELINE configs[1] at full size among
them with its own end-to-end figure).

Prints ONE JSON line (rank 0).  `--impl reference` instead times the unmodified reference's own
pthread pool (oracle/_ref/libmdzref.so) on the host cores, on a bounded sample of the same view.
"""
import argparse
import hashlib
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
PROFILE_SUMMARY = os.path.join(ROOT, "profiles", "r2_ncu_summary.json")


def macs_per_iteration(n_limbs):
    """one general product N^2 + two squarings N(N+1)/2 each."""
    return 2 * n_limbs * n_limbs + n_limbs


# name -> (view factory, text, dtype text, scaling)
def workload(name, world=1, scaling="strong"):
    from views import config2, config4, config4m
    if name == "target":
        return (config4m(3840, 2160, 100000, mode="mpfr", precision=512),
                "north-star target / BASELINE configs[3] geometry: 1e-120 wide view next to the period-707 minibrot at "
                "M(7,2), 3840x2160, MPFR 512 bits, depth 100000, 2 % of the pixels run to depth",
                "16 x u32 significand + i32 exponent (soft-float == MPFR at 512 bits, round to nearest even)", "strong")
    if name == "targetgmp":
        return (config4m(3840, 2160, 100000, mode="gmp", precision=512),
                "BASELINE configs[3]: 1e-120 wide view next to the period-707 minibrot at M(7,2), 3840x2160, GMP mpf 512 bits, "
                "depth 100000, 2 % of the pixels run to depth",
                "20 x u32 words (9 + 1 64-bit limbs) + limb exponent (soft-float == GMP 6.3 mpf at 512 bits, truncating)", "strong")
    if name == "cfg4":
        return (config4(3840, 2160, 100000, mode="mpfr", precision=512),
                "round 1's target view: 1e-120 wide on the Misiurewicz point M(23,2), 3840x2160, MPFR 512 bits, depth 100000, "
                "every pixel escapes after ~12 800 iterations",
                "16 x u32 significand + i32 exponent (soft-float == MPFR at 512 bits, round to nearest even)", "strong")
    if name == "cfg2":
        if world == 1 or scaling == "strong":
            v = config2(1920, 1080, 10000)
        else:                                   # weak: N x the pixels of the same view
            f = world ** 0.5
            v = config2(int(round(1920 * f / 8.0)) * 8, int(round(1080 * f / 8.0)) * 8, 10000)
        return (v, "BASELINE configs[1]: full M-set cx=-0.5 cy=0 size=4, %dx%d, long double mode, depth 10000"
                % (v.real_width, v.real_height),
                "u64 significand + i32 exponent (soft-float == x87 long double, round to nearest even)",
                "strong" if world == 1 else scaling)
    raise SystemExit("unknown workload " + name)


def iterations_of(raw, depth):
    import numpy as np
    return int(np.where(raw > 0, raw, depth).astype(np.int64).sum())


def rank_lines(user_height, aa, rank, world):
    """real lines of rank r's bands (r, r + world, ...): the partition mdzcuda_plan_create(first=r, stride=world) renders"""
    return [l for b in range(rank, user_height, world) for l in range(b * aa, (b + 1) * aa)]


def sample_of(lines, count):
    """`count` of the given lines, evenly spread, first and last included"""
    if count >= len(lines):
        return list(lines)
    if count <= 1:
        return [lines[len(lines) // 2]]
    return sorted(set(lines[int(round(k * (len(lines) - 1) / (count - 1)))] for k in range(count)))


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


# ---------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------
REF_SAMPLE = (192, 108)      # the bounded sample of the target view: the same rect at 1/400 of the pixels


def reference_sample_view(name):
    """The workload's own view -- same rect, precision, depth, mode -- at a resolution the host cores finish in
    seconds.  (The whole 3840x2160 frame is ~65 G pixel-iterations: a quarter of an hour on 16 cores.)"""
    from views import config2, config4, config4m
    w, h = REF_SAMPLE
    if name == "target":
        return config4m(w, h, 100000, mode="mpfr", precision=512), "the same view at %dx%d (1/400 of the pixels)" % (w, h)
    if name == "targetgmp":
        return config4m(w, h, 100000, mode="gmp", precision=512), "the same view at %dx%d (1/400 of the pixels)" % (w, h)
    if name == "cfg4":
        return config4(w, h, 100000, mode="mpfr", precision=512), "the same view at %dx%d (1/400 of the pixels)" % (w, h)
    return config2(1920, 1080, 10000), "the complete 1920x1080 frame"


def run_reference(args, emit):
    """Reference arm: the unmodified reference pool on the host cores (rank 0 only).  Maps nothing of the
    product: mdz_b200 is imported for its host-side view builder only and loads libmdzcuda lazily."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import refpath
    lib = refpath.load()
    cores = os.cpu_count() or 1
    view, sample_txt = reference_sample_view(args.workload)
    full, text, dtype, _ = workload(args.workload, 1, "strong")
    kind = "reference"
    if lib is None:
        import portpath
        kind = "port"

    def one(v):
        t0 = time.perf_counter()
        raw = refpath.ref_render(lib, v, cores)[0] if lib is not None else portpath.port_render(v, cores)
        return time.perf_counter() - t0, raw

    times, iters = [], 0
    for i in range(args.warmup + args.steps):
        dt, raw = one(view)
        if i >= args.warmup:
            times.append(dt)
            iters = iterations_of(raw, view.depth)
    total = sum(times)
    value = iters * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": text, "same_config": True,
                   "sample": "%s, %d pixel-iterations per step" % (sample_txt, iters)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%s, %d pixel-iterations per step, reference pthread pool with -t %d"
                                   % (sample_txt, iters, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.workload != "cfg2" and not args.no_side:
        # the hardware-precision configuration, complete (BASELINE configs[1]): affordable on host cores
        from views import config2
        v2 = config2(1920, 1080, 10000)
        best = None
        for _ in range(3):
            dt, raw = one(v2)
            best = dt if best is None or dt < best else best
        line["per_precision"] = {"ld64_cfg2": {"value": iterations_of(raw, v2.depth) / best, "unit": UNIT, "ms": best * 1e3,
                                               "view": "1920x1080 complete", "cores": cores}}
    emit(line)


# ---------------------------------------------------------------------------
# GPU arm, side measurements (N = 1 only)
# ---------------------------------------------------------------------------
def time_plan(torch, plan, repeats=1):
    st = torch.cuda.current_stream().cuda_stream
    plan.launch(st); plan.wait()                        # warm-up
    best = None
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.launch(st); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    return best


def side_precisions(torch, device, peak, peak32, skip=()):
    """Kernel-only it/s at the other precisions of the metric; BASELINE configs[1] (long double, complete
    1920x1080) also end to end through the C ABI with host buffers."""
    import numpy as np
    import mdz_b200
    from views import make_view, SEAHORSE, deep_embedded_julia, config2
    out = {}
    sea = lambda w, h, **kw: make_view(SEAHORSE[0], SEAHORSE[1], "1e-12", w, h, depth=10000, **kw)     # noqa: E731
    cases = [
        ("ld64_cfg2", config2(1920, 1080, 10000)),
        ("mpfr80", sea(1920, 1080, precision=80)),
        ("mpfr128", sea(1920, 1080, precision=128)),
        ("mpfr320", deep_embedded_julia(1920, 1080)),
        ("mpfr512", sea(1920, 1080, precision=512)),
        ("gmp512", sea(1920, 1080, mode="gmp", precision=512)),
        ("mpfr1024", sea(960, 540, precision=1024)),
        # 16 / 32 lanes per pixel: a few thousand pixels in flight, so the views hold at least forty times that
        ("mpfr2048", sea(480, 270, precision=2048)),
        ("mpfr4096", sea(320, 180, precision=4096)),
        ("mpfr8192", sea(240, 135, precision=8192)),
        ("gmp768", sea(960, 540, mode="gmp", precision=768)),
        ("gmp1024", sea(480, 270, mode="gmp", precision=1024)),
        ("gmp2048", sea(240, 135, mode="gmp", precision=2048)),
    ]
    stream = torch.cuda.current_stream().cuda_stream
    for name, view in cases:
        if name in skip:
            continue
        try:
            plan = mdz_b200.Plan(view, device)
        except mdz_b200.MdzCudaError as ex:
            out[name] = {"value": None, "unit": UNIT, "error": str(ex)}
            continue
        ms = time_plan(torch, plan, 3 if name == "ld64_cfg2" else 1)
        iters = iterations_of(plan.fetch(), view.depth)
        ki = plan.kernel_info()
        rate = iters / (ms * 1e-3)
        # GMP mode multiplies the top P 64-bit limbs of each operand: N = 2P words
        n_mac = ki["limbs"] - 2 if view.mode == 2 else ki["limbs"]
        out[name] = {"value": rate, "unit": UNIT, "ms": ms, "pixel_iterations": iters,
                     "view": "%dx%d" % (view.real_width, view.real_height),
                     "limbs": ki["limbs"], "mac_limbs": n_mac, "macs_per_iteration": macs_per_iteration(n_mac),
                     "lanes_per_pixel": ki.get("lanes_per_pixel", 1),
                     "regs": ki["regs_per_thread"], "spill_bytes": ki["local_bytes"],
                     "blocks_per_sm": ki["blocks_per_sm"],
                     "imad_frac": rate * macs_per_iteration(n_mac) / peak,
                     "frac_of_imad32_issue": rate * macs_per_iteration(n_mac) / peak32}
        plan.close()
        if name == "ld64_cfg2":
            res = np.empty((view.real_height, view.real_width), dtype=np.int32)
            for _ in range(2):
                p2 = mdz_b200.Plan(view, device); p2.run(res, stream); p2.close()
            t0 = time.perf_counter()
            n = 10
            for _ in range(n):
                p2 = mdz_b200.Plan(view, device); p2.run(res, stream); p2.close()
            dt = (time.perf_counter() - t0) / n
            out[name]["e2e"] = {"value": iters / dt, "unit": UNIT, "ms": dt * 1e3, "steps": n,
                                "h2d_bytes_per_step": 16 * (view.real_width + view.real_height + 2),
                                "d2h_bytes_per_step": int(res.nbytes)}
            # NOT the metric: the same render with the exact periodicity check on (what the rth_* drop-in runs)
            ms2 = []
            for _ in range(4):
                ta = time.perf_counter()
                p3 = mdz_b200.Plan(view, device)
                p3.set_cycle_detection(True)
                got = p3.run(stream=stream)
                p3.close()
                ms2.append((time.perf_counter() - ta) * 1e3)
            out[name]["cycle_detection"] = {"e2e_ms": min(ms2[1:]), "identical_raw_data": bool(np.array_equal(got, res)),
                                            "note": "opt-in exact periodicity check; no it/s figure is derived from it"}
    return out


def static_profile(kernel_key):
    """ncu figures of a kernel from the committed capture (profiles/): a static fact of the build, NOT measured in
    this run -- the JSON says so."""
    try:
        with open(PROFILE_SUMMARY) as f:
            both = json.load(f)
    except (OSError, ValueError):
        return None, None
    ent = both.get(kernel_key)
    if not ent:
        return None, both.get("_meta")
    return ent, both.get("_meta")


def ncu_num(ent, key):
    try:
        v, unit = str(ent[key]).split()[:2]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    except (KeyError, ValueError, IndexError, TypeError):
        try:
            return float(ent[key])
        except (KeyError, ValueError, TypeError):
            return None


def main():
    # rank 0 must print exactly ONE line on stdout; NCCL / torch may write banners to fd 1
    # ("NCCL version ..."), so park the real stdout and send everything else to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-side", action="store_true", help="skip the per-precision side measurements")
    ap.add_argument("--no-check", action="store_true", help="skip the untimed comparison with the reference's line driver")
    ap.add_argument("--workload", default="target", choices=["target", "targetgmp", "cfg4", "cfg2"],
                    help="target (default): the north star's 3840x2160 MPFR-512 image, split over the ranks; targetgmp: the same "
                         "in GMP mpf mode; cfg4: round 1's minibrot-free view; cfg2: BASELINE configs[1] (long double)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="cfg2 only, N>1: strong = split the 1920x1080 image (default), weak = N x the pixels")
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
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")      # barriers that do not put a spinning kernel on the GPUs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    view, wl_text, dtype_text, scaling = workload(args.workload, world, args.scaling)
    deep = args.workload != "cfg2"
    plan = mdz_b200.Plan(view, local, band_first=rank, band_stride=world)
    plan.set_order(centre_out=True)      # the queue starts with the middle bands (include/mdzcuda.h: mdzcuda_plan_set_order)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # ---- device-timed: W warm-up renders, then exactly K, barrier + synchronize on both sides ----
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
    mine = rank_lines(view.user_height, view.aa_factor, rank, world)
    my_lines = raw[mine]
    my_iters = iterations_of(my_lines, view.depth)
    t = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device="cuda")
    it = torch.tensor([my_iters, my_launches, int(mdz_b200.fallback_lines())], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
    ms_total, kernel_ms = float(t[0]), float(t[1])
    total_iters, total_launches, fallback_lines = int(it[0]), int(it[1]), int(it[2])
    value = total_iters * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the public call: host view in, host raw_data out ----
    ki = plan.kernel_info()
    xs_bytes = (ki["limbs"] + 2) * 4 * (view.real_width + plan.local_lines() + 2)
    e2e_steps = max(1, min(args.steps, 2 if deep else 20))
    out = np.empty((view.real_height, view.real_width), dtype=np.int32)   # the caller's raw_data (pageable, as MDZ's malloc)
    out.fill(-1)
    for _ in range(1 if deep else 2):                                           # untimed: pool warm, pages touched
        p2 = mdz_b200.Plan(view, local, band_first=rank, band_stride=world); p2.set_order(True); p2.run(out, stream); p2.close()
    barrier()
    e2e_ms = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        p2 = mdz_b200.Plan(view, local, band_first=rank, band_stride=world)   # prologue + H2D of the tables
        p2.set_order(True)
        p2.run(out, stream)                                                      # kernel; bands D2H as they complete
        p2.close()
        e2e_ms.append(round((time.perf_counter() - ta) * 1e3, 3))
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_iters * e2e_steps / float(te[0])
    e2e_same = bool(np.array_equal(out[mine], my_lines))

    # ---- untimed: every rank checks sampled lines of its share against the reference's own line driver ----
    cores = os.cpu_count() or 1
    check = {"identical": None, "lines": 0, "seconds": 0.0}
    ref_lib = None
    if not args.no_check:
        try:
            import refpath
            ref_lib = refpath.load()
        except Exception:
            ref_lib = None
    if ref_lib is not None:
        threads = max(1, cores // world)
        lines = sample_of(mine, threads if deep else min(len(mine), 8 * threads))
        tc = time.perf_counter()
        want = refpath.ref_render_lines(ref_lib, view, lines, threads)
        check = {"identical": bool(np.array_equal(want, raw[lines])), "lines": len(lines),
                 "seconds": time.perf_counter() - tc}
    flags = [check]
    digest = hashlib.sha256(np.ascontiguousarray(my_lines).tobytes()).hexdigest()
    digests = [digest]
    if world > 1:
        flags = [None] * world
        digests = [None] * world
        dist.all_gather_object(flags, check, group=cpu_group)
        dist.all_gather_object(digests, digest, group=cpu_group)
    host_barrier()

    # ---- untimed: the library's own multi-device call, from ONE process over all N devices ----
    single = None
    if rank == 0:
        devs = tuple(range(world))
        try:
            if world > 1:
                mdz_b200.render(view, devs)                       # contexts, pools, module load on the other devices
            ts = time.perf_counter()
            whole = mdz_b200.render(view, devs)
            dt = time.perf_counter() - ts
            same = [hashlib.sha256(np.ascontiguousarray(whole[rank_lines(view.user_height, view.aa_factor, r, world)]).tobytes()).hexdigest()
                    == digests[r] for r in range(world)]
            single = {"call": "mdzcuda_render(view, raw_host, ndev=%d)" % world, "devices": list(devs),
                      "e2e_value": iterations_of(whole, view.depth) / dt, "unit": UNIT, "ms": dt * 1e3,
                      "identical_to_per_rank_result": bool(all(same)),
                      "fallback_lines": int(mdz_b200.fallback_lines())}
        except mdz_b200.MdzCudaError as ex:
            single = {"call": "mdzcuda_render(view, raw_host, ndev=%d)" % world, "error": str(ex)}
        torch.cuda.set_device(local)
    host_barrier()

    if rank == 0:
        peak = mdz_b200.imad_peak(local, 200)                  # IMAD.WIDE.U32.X chains: 32x32->64 MAC/s
        peak32 = mdz_b200.imad_peak(local, 200, wide=False)    # 32-bit IMAD issue rate
        macs = macs_per_iteration(ki["limbs"] - 2 if view.mode == 2 else ki["limbs"])     # GMP mode multiplies the top P 64-bit limbs: N = 2P words
        kernel_rate = (total_iters / world) / (kernel_ms * 1e-3)        # this GPU's kernel alone
        kernel_key = {"target": "mpfr512_target", "targetgmp": "gmp512_target", "cfg4": "mpfr512_target", "cfg2": "ld64_cfg2_phase0"}[args.workload]
        prof, meta = static_profile(kernel_key)
        dram = None
        if prof:
            parts = [ncu_num(prof, k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
            dram = sum(parts) if all(x is not None for x in parts) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": dtype_text, "data": "synthetic",
            "config": {"workload": wl_text + ("; one image split over %d GPU(s)" % world),
                       "pixel_iterations_per_step": total_iters,
                       "partition": "interleaved line bands, one plan per rank, no collective; each plan's pixel queue starts with its middle bands",
                       "l2": "256 MiB buffer rewritten between steps (inputs are KB-sized tables)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                    "h2d_bytes_per_step": xs_bytes, "d2h_bytes_per_step": int(my_lines.nbytes),
                    "ms_each": e2e_ms, "identical_to_device_timed_result": e2e_same,
                    "includes": "mdzcuda_plan_create (host prologue with libmpfr/long double, H2D of the tables) + "
                                "mdzcuda_plan_run (kernel, D2H of finished bands into pageable host raw_data) + destroy"},
            "gpu_launches": total_launches,
            "fallback_lines": fallback_lines,
            "clocks": clocks,
            "identical_to_reference_on_sample": (all(f["identical"] for f in flags) if all(f["identical"] is not None for f in flags) else None),
            "reference_check": {"per_rank": flags,
                                "how": "each rank: the reference's own line driver (oracle/_ref/libmdzref.so, ref_render_lines) on evenly "
                                       "spread lines of its share of the full-size view, compared with the device-timed result; untimed"},
            "single_process_ndev": single,
            "roofline": {"bound": "imad", "achieved": kernel_rate * macs / 1e12, "peak": peak / 1e12,
                         "unit": "T 32x32->64 MAC/s", "frac": kernel_rate * macs / peak,
                         "peak_imad32": peak32 / 1e12, "frac_of_imad32_issue": kernel_rate * macs / peak32,
                         "traffic": dram,
                         "traffic_note": None if dram is None else
                             "dram__bytes_read + dram__bytes_write of ONE launch of this kernel in the committed ncu --set full capture "
                             "(profiles/r2_ncu_summary.json: %s), which renders the bench view at 960x540 -- 1/16 of the bench frame's pixels: "
                             "algorithmic bytes of that launch are 2.07 MB of raw_data out + 0.1 MB of tables in; the results still sit in "
                             "L2 when the kernel ends, hence less than that reaches DRAM.  Static figure, not measured in this run" % kernel_key,
                         "binding_pipe": None if not prof else {
                             "static": True, "source": "profiles/r2_ncu_summary.json: " + kernel_key,
                             "captured_at_rev": (meta or {}).get("rev"),
                             "alu_pct": ncu_num(prof, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                             "fma_pct": ncu_num(prof, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                             "fmaheavy_cycles_pct": ncu_num(prof, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                             "fp64_pct": ncu_num(prof, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                             "issue_active_pct": ncu_num(prof, "smsp__issue_active.avg.pct_of_peak_sustained_active")},
                         "note": "integer-pipe roofline (SURVEY 8d), not hbm/tensor. achieved = this GPU's it/s x W(N)=2N^2+N "
                                 "full-schoolbook MACs (N=%d limbs -> %d; the kernel forms only the high ~59%% of each "
                                 "product, DESIGN.md 2.1); peak = IMAD.WIDE.U32.X carry-chain microbenchmark measured in "
                                 "this run (MEASURED_PEAKS.json has no integer peak); peak_imad32 = 32-bit IMAD issue rate, "
                                 "twice that; kernel avg %.3f ms per render by CUDA events on the launching stream (%d launch(es) per render); "
                                 "HBM traffic is 4 B per pixel out; traffic / binding_pipe are static figures of the committed ncu capture"
                                 % (ki["limbs"], macs, kernel_ms, total_launches // max(1, args.steps * world))},
            "kernel": ki,
        }
        # CPU baseline: the unmodified reference pool on this box's cores, bounded sample of the same view
        try:
            if ref_lib is None:
                raise RuntimeError("oracle/_ref/libmdzref.so not available")
            sview, sample_txt = reference_sample_view(args.workload)
            tb = time.perf_counter()
            rraw, _ = refpath.ref_render(ref_lib, sview, cores)
            dt = time.perf_counter() - tb
            line["cpu_baseline"] = {"value": iterations_of(rraw, sview.depth) / dt, "unit": UNIT,
                                    "cores": cores, "kind": "reference",
                                    "sample": "%s, unmodified reference pool, -t %d, %.2f s" % (sample_txt, cores, dt)}
        except Exception as ex:   # the reference .so is optional on the box
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
        if not args.no_side and world == 1:
            line["per_precision"] = side_precisions(torch, local, peak, peak32,
                                                    skip=("ld64_cfg2",) if args.workload == "cfg2" else ())
        emit(line)
    plan.close()
    host_barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
