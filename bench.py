#!/usr/bin/env python
"""bench.py -- CPHF polarizability hot path (AO2MO + Hessian assembly + multi-frequency solve).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # reference algorithm on host cores

Metric (BASELINE.json): FP64 TFLOP/s of the end-to-end CPHF property calculation, counted with
the algorithmic flop model of SURVEY.md section 8(d) (never the flops actually executed), plus
wall time per step (``ms_per_step``).  One step = one full property calculation:
AO->MO transform of the resident AO ERI tensor, A/B assembly, every frequency solved, property
tensors formed.

Workloads (``config.workload``):
  N = 1, 2, 4 : BASELINE config 5 -- synthetic RHF nbasis=256, nocc=32: magnetizability (static,
                imaginary operator) + 64-frequency polarizability sweep.  It is the largest config
                whose AO ERI tensor (34.4 GB) fits one B200; config 3 (nbasis=512: 550 GB) needs 8.
                For N > 1 the same job is strong-scaled: I sharded over nu, frequencies dealt out.
  N = 8       : BASELINE config 3 -- synthetic RHF nbasis=512, nocc=64, polarizability at 8
                frequencies, AO ERI tensor sharded 8 ways (68.7 GB per GPU).
Inputs are synthetic (seeded), generated on the device outside the timed region.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

METRIC = "cphf_polarizability_fp64_tflops"
UNIT = "TFLOP/s"


# ----------------------------------------------------------------------------------- workloads
def workload_for(n_gpus: int, args):
    if args.n:
        n, o = args.n, args.nocc or max(1, args.n // 8)
        nf = args.nfreq or 8
        return {"name": f"custom synthetic RHF nbasis={n} nocc={o}, polarizability at {nf} frequencies",
                "n": n, "o": o, "pol_freqs": list(np.linspace(0.0, 0.35, nf)) if nf > 1 else [0.0],
                "magnetizability": bool(args.mag)}
    if n_gpus >= 8:
        return {"name": "config 3: synthetic RHF nbasis=512 nocc=64, polarizability at 8 frequencies, "
                        "AO ERIs sharded over 8 GPUs",
                "n": 512, "o": 64, "pol_freqs": list(np.linspace(0.0, 0.35, 8)), "magnetizability": False}
    return {"name": "config 5: synthetic RHF nbasis=256 nocc=32, magnetizability (static, imaginary) + "
                    "64-frequency polarizability sweep",
            "n": 256, "o": 32, "pol_freqs": list(np.linspace(0.0, 0.63, 64)), "magnetizability": True}


def algorithmic_flops(w) -> float:
    from pymolresponse_b200 import synthetic

    n, o = w["n"], w["o"]
    nov = o * (n - o)
    freqs = w["pol_freqs"]
    n_static = sum(1 for f in freqs if f == 0.0)
    n_dyn = len(freqs) - n_static
    f = synthetic.flops_ao2mo_rhf_partial(n, o)
    f += synthetic.flops_solve_rhf(nov, n_static, n_dyn, 3)
    if w["magnetizability"]:
        # static K-solve of the imaginary operator; potrf(K) is shared with the sweep when there
        # is one, otherwise it costs one u
        f += 4.0 * nov * nov * 3
        if n_dyn == 0:
            f += nov**3 / 3.0
    return f


# ----------------------------------------------------------------------------------- helpers
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1])); mx.append(float(t[2])); pw.append(float(t[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), t[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            hi = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
            out = {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(pw)}
        return out


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def fp64_dgemm_peak(torch):
    """The roofline denominator for the FP64 kernels: cuBLAS DGEMM 8192^3 on this GPU, best of 5
    (burst), measured in this run -- MEASURED_PEAKS.json carries bf16 and HBM only."""
    nn = 8192
    a = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
    b = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * nn**3 / (best * 1e-3) / 1e12


def build_inputs(w, rank, world, torch):
    """Seeded synthetic inputs; the AO ERI slab I[:, nu_lo:nu_hi] is built on the device from the
    factor B (input generation, cuBLAS, outside the timed region)."""
    from pymolresponse_b200 import distributed as pd
    from pymolresponse_b200 import synthetic

    n, o = w["n"], w["o"]
    B = synthetic.eri_factor(n)
    lo, hi = pd.shard_range(n, rank, world)
    cnt = hi - lo
    Bd = torch.from_numpy(B).cuda()                          # (naux, n, n)
    Bf = Bd.reshape(B.shape[0], n * n)
    I = torch.empty((n, cnt, n, n), dtype=torch.float64, device="cuda")
    rows_per = max(1, (1 << 28) // (n * n))                  # <= 2 GiB of output per GEMM
    for mu in range(n):
        left = Bd[:, mu, lo:hi]                              # (naux, cnt)
        for r0 in range(0, cnt, rows_per):
            r1 = min(cnt, r0 + rows_per)
            I[mu, r0:r1] = (left[:, r0:r1].t() @ Bf).reshape(r1 - r0, n, n)
    del Bd, Bf
    torch.cuda.empty_cache()
    if world == 1:
        # make the 8-fold permutational symmetry exact to the last bit (cuBLAS rounds (mu nu| and
        # (nu mu| independently), so that the resident and the end-to-end paths see the same tensor
        from pymolresponse_b200 import _native as nv

        nv.check(nv.lib().pmr_eri_complete_s8(nv.ptr(I), n, 0, nv.stream_ptr()))
        torch.cuda.synchronize()
    else:
        from pymolresponse_b200 import _native as nv

        nv.check(nv.lib().pmr_eri_complete_ket_s2(nv.ptr(I), n * cnt, n, nv.stream_ptr()))
        torch.cuda.synchronize()
    C = synthetic.mo_coefficients(n, synthetic.SEED_ALPHA)[None]
    E = np.diag(synthetic.orbital_energies(n, o, synthetic.SEED_ALPHA))[None]
    return {"I": I, "nu_range": (lo, hi), "C": C, "E": E, "occ": (o, n - o, o, n - o),
            "O_real": synthetic.operator_integrals(n, 3, False),
            "O_imag": synthetic.operator_integrals(n, 3, True), "B": B}


def one_step(w, inp, I, device_resident: bool, world: int, stage_ms: dict | None = None,
             ao_symmetry: str = "s1"):
    """One full property calculation through the public API.  ``I`` is a CUDA tensor (kernel
    timing, inputs resident) or a host NumPy array (end-to-end timing, H2D inside)."""
    from pymolresponse_b200 import cphf, distributed as pd, integrals, solvers
    from pymolresponse_b200.ao2mo import AO2MO
    from pymolresponse_b200.core import Hamiltonian, Program, Spin
    from pymolresponse_b200.properties.electric import Polarizability
    from pymolresponse_b200.properties.magnetic import Magnetizability

    import torch

    marks = []

    def mark(name):
        if stage_ms is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    mark("start")
    nu_range = inp["nu_range"] if world > 1 else None
    a = AO2MO(inp["C"], inp["occ"], I=I, device_results=True, nu_range=nu_range, reduce=world > 1,
              ao_symmetry=ao_symmetry)
    a.perform_rhf_partial()
    tei = a.tei_mo
    mark("ao2mo")
    gen = integrals.IntegralsArrays({"dipole": inp["O_real"], "angmom_common_gauge": inp["O_imag"]})
    out = {}
    s_mag = None
    if w["magnetizability"]:
        s = s_mag = solvers.ExactInvCholesky(inp["C"], inp["E"], inp["occ"])
        s.device_results = device_resident
        s.tei_mo = tei
        mag = Magnetizability(Program.Arrays, gen, cphf.CPHF(s))
        mag.form_operators()
        mag.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet)
        mag.form_results()
        out["magnetizability"] = mag.magnetizability
        mark("magnetizability")
    s = solvers.ExactInv(inp["C"], inp["E"], inp["occ"])
    s.device_results = device_resident
    s.distribute_frequencies = world > 1
    s.tei_mo = tei
    if s_mag is not None:
        s_mag.share_hessian_with(s)   # same reference, Hamiltonian and spin: M, K, chol(K) built once
    pol = Polarizability(Program.Arrays, gen, cphf.CPHF(s), frequencies=[float(f) for f in w["pol_freqs"]])
    pol.form_operators()
    pol.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet)
    pol.form_results()
    out["polarizabilities"] = pol.polarizabilities
    out["stats"] = getattr(s, "solve_stats", {})
    mark("polarizability_sweep")
    if stage_ms is not None:
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            stage_ms[name] = e0.elapsed_time(e1)
    return out


# ----------------------------------------------------------------------------------- CPU baseline
def cpu_sample_spec():
    return {"n": 128, "o": 16, "freqs": [0.0, 0.21, 0.42, 0.63]}


def cpu_reference_step(spec=None):
    """The reference's algorithm (oracle port: np.dot + einsum quarter transforms, explicit
    super-Hessian and np.linalg.inv per frequency) on a bounded sample of the workload."""
    from oracle import pmr_oracle as orc
    from pymolresponse_b200 import synthetic

    spec = spec or cpu_sample_spec()
    n, o = spec["n"], spec["o"]
    case = synthetic.rhf_case(n, o)
    Or = synthetic.operator_integrals(n, 3, False)
    t0 = time.perf_counter()
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    t1 = time.perf_counter()
    out = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial",
                       [{"ao_integrals": Or, "is_imaginary": False}], spec["freqs"])
    t2 = time.perf_counter()
    nov = o * (n - o)
    n_static = sum(1 for f in spec["freqs"] if f == 0.0)
    flops = synthetic.flops_ao2mo_rhf_partial(n, o) + synthetic.flops_solve_rhf(
        nov, n_static, len(spec["freqs"]) - n_static, 3)
    return {"seconds": t2 - t0, "ao2mo_s": t1 - t0, "solve_s": t2 - t1, "flops": flops,
            "alpha_xx": float(out["results"][0][0, 0])}


def sample_text(spec):
    return (f"synthetic RHF nbasis={spec['n']} nocc={spec['o']}, polarizability at "
            f"{len(spec['freqs'])} frequencies (reference algorithm: einsum AO2MO, explicit (2nov)^2 "
            "Hessian, np.linalg.inv per frequency); TFLOP/s uses the same algorithmic flop model")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = cpu_sample_spec()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_step(spec)
    times, flops = [], 0.0
    for _ in range(args.steps):
        r = cpu_reference_step(spec)
        times.append(r["seconds"])
        flops = r["flops"]
    t = sum(times) / len(times)
    val = flops / t / 1e12
    cores = os.cpu_count() or 1
    w = workload_for(args.gpus, args)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus < 8 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "sample": sample_text(spec)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample_text(spec)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch

    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import distributed as pd

    rank, world, local = pd.init_from_env()
    torch.cuda.set_device(local)
    w = workload_for(world, args)
    n, o = w["n"], w["o"]
    nov = o * (n - o)
    F = algorithmic_flops(w)
    lib = nv.lib()

    peak64 = fp64_dgemm_peak(torch) if rank == 0 else None
    inp = build_inputs(w, rank, world, torch)
    I_dev = inp["I"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            pd.barrier()
            torch.cuda.synchronize()

    # ---------------- kernel-only timing: inputs resident in HBM -----------------------------
    for _ in range(args.warmup):
        res = one_step(w, inp, I_dev, True, world)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nv.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = one_step(w, inp, I_dev, True, world)
    e1.record()
    sync_all()
    launches = nv.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        pd.dist().all_reduce(t, op=pd.dist().ReduceOp.MAX)
    ms = float(t.item())

    # ---------------- roofline pass: per-launch CUDA events on the dominant kernel -----------
    stages = {}
    from pymolresponse_b200 import _engine as eng
    from pymolresponse_b200 import ao2mo as ao2mo_mod

    eng.STAGE_TIMES = True          # untimed extra step: per-stage CUDA events (adds synchronisation)
    ao2mo_mod.LAST_STAGE.clear()
    res_s = one_step(w, inp, I_dev, True, world, stage_ms=stages)
    eng.STAGE_TIMES = False
    stages["ao2mo_parts"] = dict(ao2mo_mod.LAST_STAGE.get("ms", {}))
    stages["sweep_parts"] = dict((res_s.get("stats") or {}).get("ms", {}))
    sync_all()

    roofline = None
    if rank == 0:
        lib.pmr_profile_enable(1)
        lib.pmr_set_lookahead(0)   # serialise the Cholesky so every launch is timed on its own
    one_step(w, inp, I_dev, True, world)   # every rank takes part (collectives), rank 0 records
    sync_all()
    if rank == 0:
        lib.pmr_set_lookahead(1)
        import ctypes as C

        buf = (C.c_double * 24)()
        lib.pmr_profile_fetch(buf, 6)
        lib.pmr_profile_enable(0)
        prof = {}
        names = ["gemm128x128", "gemm128x128_lower(trailing update)", "gemm128x64", "gemm128x64_lower",
                 "gemm128x32", "gemm128x32_lower"]
        for c in range(6):
            cnt, tms, fl, by = buf[c * 4], buf[c * 4 + 1], buf[c * 4 + 2], buf[c * 4 + 3]
            if cnt > 0:
                prof[names[c]] = {"launches": int(cnt), "ms": tms, "tflops": fl / (tms * 1e-3) / 1e12,
                                  "algorithmic_flops": fl}
        if prof:
            top = max(prof, key=lambda k: prof[k]["ms"])
            a = prof[top]
            roofline = {"bound": "tensor", "kernel": f"dgemm_dmma_kernel [{top}]",
                        "achieved": a["tflops"], "peak": peak64, "unit": "TFLOP/s",
                        "frac": a["tflops"] / peak64 if peak64 else None, "traffic": None,
                        "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured in this run, "
                                       "burst; nominal B200 FP64 is 37-40 TFLOP/s",
                        "share_of_step": a["ms"] / ms, "avg_launch_ms": a["ms"] / a["launches"],
                        "classes": prof}

    # ---------------- end-to-end: host buffers, H2D + D2H inside the timed region ------------
    # (after the roofline pass: the resident device copy of the AO tensor is released first so the
    #  per-step H2D copy has room at config 3, 68.7 GB per GPU)
    e2e = None
    e2e_note = None
    if not args.no_e2e:
        need = I_dev.numel() * 8
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = None
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        ok_local = avail is None or avail > 1.15 * need * local_world + (8 << 30)
        flag = torch.tensor([1 if ok_local else 0], dtype=torch.int64, device="cuda")
        if world > 1:
            pd.dist().all_reduce(flag, op=pd.dist().ReduceOp.MIN)
        if int(flag.item()) == 0:
            e2e_note = (f"skipped: host RAM available {avail / 2**30:.0f} GiB < {local_world} ranks x "
                        f"{need / 2**30:.0f} GiB AO-tensor slabs")
    if not args.no_e2e and e2e_note is None:
        try:
            I_host_t = torch.empty(I_dev.shape, dtype=torch.float64, pin_memory=True)
        except Exception:
            I_host_t = torch.empty(I_dev.shape, dtype=torch.float64)
        I_host_t.copy_(I_dev)
        torch.cuda.synchronize()
        I_host = I_host_t.numpy()
        del I_dev
        inp["I"] = None
        torch.cuda.empty_cache()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        small = sum(inp[k].nbytes for k in ("C", "E", "O_real", "O_imag"))
        nfreq = len(w["pol_freqs"])
        d2h = nfreq * (3 * 2 * nov * 8 + 72) + (3 * 2 * nov * 8 + 72 if w["magnetizability"] else 0)

        e2e_diffs = []

        def timed_e2e(sym):
            one_step(w, inp, I_host, False, world, ao_symmetry=sym)
            sync_all()
            e0.record()
            for _ in range(e2e_steps):
                r = one_step(w, inp, I_host, False, world, ao_symmetry=sym)
            e1.record()
            sync_all()
            tt = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device="cuda")
            if world > 1:
                pd.dist().all_reduce(tt, op=pd.dist().ReduceOp.MAX)
            # same kernels, same tensor in HBM: the paths agree bit for bit in every run recorded so
            # far; the hard limit is the parity bar of the property tensors (1e-8)
            diff = max([float(np.abs(np.asarray(a) - np.asarray(b)).max())
                        for a, b in zip(r["polarizabilities"], res["polarizabilities"])
                        if a is not None and b is not None] or [0.0])
            assert diff <= 1e-8, f"end-to-end and resident results differ by {diff}"
            e2e_diffs.append(diff)
            return float(tt.item())

        # whole-array upload: nothing assumed about the AO tensor
        ms_full = timed_e2e("s1")
        full = {"value": F / (ms_full * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": ms_full,
                "h2d_bytes_per_step": int(I_host.nbytes * world + small),
                "d2h_bytes_per_step": int(d2h), "pinned_host_input": bool(I_host_t.is_pinned()),
                "upload": "whole AO ERI array (ao_symmetry='s1')"}
        if world > 1:
            # nu-sharded slabs: only (la si| = (si la| applies inside a slab, half of the bytes
            ms_e = timed_e2e("s8")
            e2e = {"value": F / (ms_e * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": ms_e,
                   "h2d_bytes_per_step": int(world * lib.pmr_eri_ket_s2_bytes(
                       I_host.shape[0] * I_host.shape[1], n)) + int(small),
                   "d2h_bytes_per_step": int(d2h), "pinned_host_input": bool(I_host_t.is_pinned()),
                   "upload": "ket-symmetric half (si <= la) of every rank's nu-slab "
                             "(ao_symmetry='s8'), completed on the device",
                   "whole_array_upload": full}
        elif world == 1:
            # the public API's ao_symmetry='s8': only the permutationally unique part of the AO ERI
            # tensor crosses PCIe, the rest is rebuilt in HBM (pmr_eri_upload_s8/_complete_s8)
            ms_e = timed_e2e("s8")
            e2e = {"value": F / (ms_e * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": ms_e,
                   "h2d_bytes_per_step": int(lib.pmr_eri_s8_bytes(n, nv.ERI_S8_LAYOUT)) + int(small),
                   "d2h_bytes_per_step": int(d2h), "pinned_host_input": bool(I_host_t.is_pinned()),
                   "upload": "permutationally unique part of the AO ERI array (ao_symmetry='s8': "
                             "rows (la, si <= la), columns from (la, 0) on -- a sixth of the tensor "
                             "in long runs), completed on the device",
                   "whole_array_upload": full}
        del I_host_t, I_host
        e2e["max_abs_diff_vs_resident"] = max(e2e_diffs) if e2e_diffs else None
    elif e2e_note is not None:
        e2e = {"value": None, "unit": UNIT, "note": e2e_note, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": 0}

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = None
        if not args.no_cpu:
            spec = cpu_sample_spec()
            r = cpu_reference_step(spec)
            cpu = {"value": r["flops"] / r["seconds"] / 1e12, "unit": UNIT, "cores": cores,
                   "kind": "port", "sample": sample_text(spec), "seconds": r["seconds"]}
        peaks, src = measured_peaks()
        alpha0 = res["polarizabilities"][0]
        line = {
            "metric": METRIC, "value": F / (ms * 1e-3) / 1e12, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if world < 8 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["name"], "nbasis": n, "nocc": o, "nov": nov,
                       "frequencies": len(w["pol_freqs"]), "algorithmic_flops_per_step": F,
                       "l2_policy": "inputs larger than L2 (AO ERI tensor >= 34 GB per GPU)",
                       "fp64_peak_tflops": peak64, "hbm_peak_gbs": peaks.get("hbm_gbs"),
                       # whole job against N x the measured DGEMM rate (north_star: "fraction of the
                       # FP64 roofline"); roofline.frac below is the dominant kernel alone
                       "fraction_of_fp64_peak": (F / (ms * 1e-3) / 1e12) / (world * peak64)
                       if peak64 else None,
                       "peaks_source": src},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches // max(1, args.steps)),
            "roofline": roofline, "cpu_baseline": cpu, "stage_ms": stages,
            "check": {"alpha_static_diag": [float(x) for x in np.diag(alpha0)],
                      "solve_stats": res.get("stats")},
        }
        emit(line)
    if world > 1:
        pd.barrier()
        pd.dist().destroy_process_group()


def _claim_stdout():
    """Route everything libraries print on fd 1 (NCCL banners ...) to stderr and keep a private
    handle on the real stdout, so rank 0 prints exactly ONE JSON line there."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    _REAL_STDOUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="override nbasis (experiments)")
    ap.add_argument("--nocc", type=int, default=0)
    ap.add_argument("--nfreq", type=int, default=0)
    ap.add_argument("--mag", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
