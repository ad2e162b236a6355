#!/usr/bin/env python
"""bench.py -- elements assembled/s of the hot path (BASELINE.json metric) on N B200s of one node.

A "step" = one pass of the numeric assembly (element quadrature loop fused with the deterministic
scatter into nzval) over the whole workload, with the mesh / scatter maps already resident in HBM.
`e2e` = the same metric through the reference-facing C-ABI call sequence with HOST buffers
(efg_set_mesh/_space -> efg_start -> efg_assemble (symbolic + numeric) -> efg_fetch_csc), i.e.
host->device and device->host copies and the symbolic phase inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload heat_t6|heat_q4|heat_t3|elasticity_t6|stokes_gen]
                  [--n SIZE] [--impl reference]

N > 1: launched by torchrun, one rank per GPU; the matrix columns are split between the ranks (owner computes,
halo elements replicated, no data-path collective); NCCL only carries the barrier, the max-over-ranks
time and the validation reductions.

Every run also carries a `config5` record: BASELINE config 5 (heat FEH1_Q4, 16384 x 16384 cells = 268 M elements,
nnz 2 416 017 409) SPLIT over the N ranks (strong scaling; N = 1: the whole problem on one GPU), numeric time as the
max over ranks, validated over NCCL against grid-derived nnz / pattern checksum / sum of squares.  The N = 1 run adds
`other_configs` (configs 3, 4, 5-share and 1 timed the same way as the headline workload).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (default N, description)
    "heat_t6": (4000, "heat Poisson FEH1_T6, unit-square NxN T6block (BASELINE config 2)"),
    "heat_t3": (100, "heat Poisson FEH1_T3, unit-square NxN T3block (BASELINE config 1)"),
    "heat_q4": (5792, "heat Poisson FEH1_Q4 Gauss order 2, NxN Q4block (BASELINE config 5 per-GPU share)"),
    "elasticity_t6": (2000, "plane-stress elasticity vector FEH1_T6, NxN T6block (BASELINE config 3)"),
    "stokes_gen": (1000, "Stokes Taylor-Hood T6/T3 'gen' 3-block assembly (BASELINE config 4)"),
}


def make_problem(efg, workload, n):
    if workload == "heat_t6":
        return efg.heat_problem(efg.T6, n)
    if workload == "heat_t3":
        return efg.heat_problem(efg.T3, n)
    if workload == "heat_q4":
        return efg.heat_problem(efg.Q4, n)
    if workload == "elasticity_t6":
        return efg.elasticity_problem(n, efg.T6)
    if workload == "stokes_gen":
        return efg.stokes_problem(n, "gen")
    raise SystemExit(f"unknown workload {workload}")


def triplets_per_element(prob):
    nd = {1: 1, 2: 4}
    k = prob.meshes[0].kind
    fid = prob.form.form_id
    if fid == 1:
        return k * k
    if fid == 2:
        return 4 * k * k
    if fid in (3, 5, 4):
        return 216
    return 144


def algorithmic_bytes(prob, nnz):
    """every input read once + slot map once + every output value written once, int32 device
    indices: 4*nen*nel + 16*nnodes + 4*ndofs + 4*ntriplets + 8*nnz (pressure mesh connectivity added for Stokes)."""
    b = 0
    for m in prob.meshes:
        b += 4 * m.kind * m.nel_
    b += 16 * prob.meshes[0].nnodes_
    b += 4 * getattr(prob, "ndofs_local_", prob.ndofs)     # dof map entries of this rank's sub-mesh
    b += 4 * triplets_per_element(prob) * prob.meshes[0].nel_
    b += 8 * nnz
    return b



def bench_widened_rows(torch, efg, _lib, local, steps, warmup, peak):
    """SURVEY 8f rows f4 / f5 measured like the other configs (numeric phase, CUDA events on the library's stream):
    f4: mesh + EBC + numbering of config 2 made on the device through the C ABI (efg_gen_*) and assembled from there;
    f5: the Reddy loop on the bubble / L2-pressure pairs (tiled path) and the tetrahedral heat form (two-pass path)."""
    recs = []
    # ---- f4: config 2 generated on the device ----------------------------------------------------------------------
    try:
        N = 4000
        eng = efg.Engine(local)
        stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
        eng.synchronize()
        t0 = time.perf_counter()
        eng.gen_mesh(0, efg.T6, N, N, 1.0, 1.0)
        eng.gen_space(0, 0, 1)
        tol = 1.0 / N / 100
        for box in ((0, 1, -tol, tol), (0, 1, 1 - tol, 1 + tol), (-tol, tol, 0, 1), (1 - tol, 1 + tol, 0, 1)):
            eng.setebc_box(0, 1, box[0] - tol, box[1] + tol, box[2], box[3])
        nfree, nd = eng.number_dofs([0])
        eng.synchronize()
        gen_ms = 1e3 * (time.perf_counter() - t0)
        eng.start(nd, nd)
        nnz = eng.symbolic(_lib.FORM_HEAT, 3)
        ms = time_numeric(torch, eng, stream, [1.0], steps, warmup, False) / steps
        nel = 2 * N * N
        recs.append({"row": "f4", "workload": f"heat FEH1_T6 N={N}: T6block + EBC boxes + numberdofs! made on the device (efg_gen_*), nothing crosses PCIe",
                     "generation_ms": gen_ms, "elements": nel, "ndofs": int(nd), "nunknowns": int(nfree), "nnz": int(nnz),
                     "nnz_closed_form_ok": int(nnz) == 46 * N * N + 16 * N + 1, "ms_per_step": ms, "value": nel / (ms / 1e3), "unit": "elements/s"})
        eng.close()
    except Exception as e:      # noqa: BLE001 -- side records never take the headline line down
        recs.append({"row": "f4", "error": f"{type(e).__name__}: {e}"})
    # ---- f5 -----------------------------------------------------------------------------------------------------------
    def one(name, prob, row="f5"):
        eng = efg.Engine(local)
        stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
        efg.load_problem(eng, prob)
        nnz = eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
        m = prob.meshes[0]
        nen = m.conn.shape[1]
        ndof_entries = sum((0 if sp.field is None else sp.field.dofnums.size) + (0 if sp.cellfield is None else sp.cellfield.dofnums.size)
                           for sp in prob.spaces)
        nv = prob.spaces[0].fe.nbf                 # COO triplets the reference appends per element (src/Assemblers.jl:97-114)
        ntrip = 4 * nv * nv + 4 * nv * prob.spaces[2].fe.nbf if prob.form.form_id == _lib.FORM_STOKES_REDDY else nv * nv
        alg = 4 * nen * m.nel + 8 * m.xy.shape[1] * m.nnodes + 4 * ndof_entries + 4 * ntrip * m.nel + 8 * nnz
        small = alg < 2 * L2_BYTES
        ms = time_numeric(torch, eng, stream, prob.form.params(), steps, warmup, small) / steps
        path = int(eng.stat(_lib.STAT_PATH))
        rec = {"row": row, "workload": name, "elements": m.nel, "ndofs": int(prob.ndofs), "nnz": int(nnz), "path": {1: "two-pass", 2: "tiled-fused"}[path],
               "ms_per_step": ms, "value": m.nel / (ms / 1e3), "unit": "elements/s", "algorithmic_bytes_per_element": alg / m.nel,
               "roofline": {"achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "frac": alg / (ms / 1e3) / 1e9 / peak, "unit": "GB/s"}}
        eng.close()
        return rec
    for name, make in (("Stokes Reddy FEH1_T3_BUBBLE / FEH1_T3 (p1b_p1.jl), N=1200", lambda: efg.stokes_f5_problem(1200, "p1b_p1")),
                       ("Stokes Reddy FEH1_Q4 / FEL2_Q4 (q1_q0.jl), N=2000", lambda: efg.stokes_f5_problem(2000, "q1_q0")),
                       ("heat FEH1_T4 (t4.jl), 100^3 cells = 6 M tetrahedra", lambda: efg.heat_problem(efg.T4, 100))):
        try:
            recs.append(one(name, make()))
        except Exception as e:      # noqa: BLE001
            recs.append({"row": "f5", "workload": name, "error": f"{type(e).__name__}: {e}"})
    return recs


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(efg, workload, sample_n):
    """The oracle (C restatement of the reference's serial COO + sparse() path) on one host core."""
    from oracle import oracle as orc
    orc.build()
    prob = make_problem(efg, workload, sample_n)
    tm = {}
    t0 = time.perf_counter()
    _, _, nz = orc.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs, timing=tm)
    dt = time.perf_counter() - t0
    return {"value": prob.nel / dt, "unit": "elements/s", "cores": 1, "kind": "port",
            "sample": f"{workload} N={sample_n}: {prob.nel} elements, integrate {tm['integrate_s']:.2f} s + finish(sparse) "
                      f"{tm['finish_s']:.2f} s; C restatement of Elfel's CPU path (gcc -O2, no FMA), not Julia"}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; Julia is absent) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import elfel_jl_b200 as efg
    from oracle import oracle as orc
    orc.build()
    n = args.ref_n
    prob = make_problem(efg, args.workload, n)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    v = prob.nel / (ms / 1e3)
    sample = f"{args.workload} N={n} ({prob.nel} elements) per step; C restatement of Elfel's serial COO + sparse() path, 1 thread"
    print(json.dumps({
        "impl": "reference", "metric": "elements assembled/s", "value": v, "unit": "elements/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOADS[args.workload][1]}, bounded sample N={n}"},
        "cpu_baseline": {"value": v, "unit": "elements/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


L2_BYTES = 126 << 20


def load_peak():
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        return json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def time_numeric(torch, eng, stream, params, steps, warmup, flush_l2):
    """K numeric launches timed with CUDA events on the library's stream.  Working sets larger than L2 are timed back to
    back (one event pair around the K launches); smaller ones are timed launch by launch with a write of a buffer
    twice the size of L2 in between (outside the event pairs).  -> total ms of the K launches."""
    for _ in range(warmup):
        eng.numeric(params)
    eng.synchronize()
    if not flush_l2:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(steps):
            eng.numeric(params)
        ev1.record(stream)
        ev1.synchronize()
        return ev0.elapsed_time(ev1)
    junk = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device="cuda")
    pairs = []
    with torch.cuda.stream(stream):
        for _ in range(steps):
            junk.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.numeric(params)
            e1.record(stream)
            pairs.append((e0, e1))
    eng.synchronize()
    torch.cuda.synchronize()
    return float(sum(a.elapsed_time(b) for a, b in pairs))


def bench_device_problem(torch, efg, _lib, local, workload, n, steps, warmup, peak):
    """One of the `other_configs`: mesh and numbering generated on the GPU (elfel.jl_b200/sharding.py; Stokes: host mirror),
    handed to the library as device pointers, numeric phase timed like the headline workload."""
    from elfel_jl_b200.sharding import shard_problem
    t0 = time.perf_counter()
    if workload == "stokes_gen":
        prob = make_problem(efg, workload, n)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        meshes = [(m.kind, dev(m.conn), dev(m.xy)) for m in prob.meshes]
        dofs = [dev(sp.field.dofnums) for sp in prob.spaces]
    else:
        prob, _, _ = shard_problem(efg, workload, n, 0, 1)
        meshes = [(m.kind, m.conn, m.xy) for m in prob.meshes]
        dofs = [sp.field.dofnums for sp in prob.spaces]
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    for m, (_, c, x) in zip(prob.meshes, meshes):
        m.nel_, m.nnodes_ = int(c.shape[0]), int(x.shape[0])
    prob.ndofs_local_ = int(sum(d.numel() for d in dofs))
    eng = efg.Engine(local)
    te = int(os.environ.get("EFG_BENCH_TE_" + workload.upper(), 0))      # kernel tuning (tools/ab2.py): forced tile size
    if te:
        eng.set_option(_lib.OPT_TILE_ELEMS, te)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
    for slot, (kind, c, x) in enumerate(meshes):
        eng.set_mesh(slot, kind, c, x)
    for slot, (d, ms) in enumerate(zip(dofs, prob.space_mesh)):
        eng.set_space(slot, ms, d)
    eng.start(prob.ndofs, prob.ndofs)
    t0 = time.perf_counter()
    nnz = eng.symbolic(prob.form.form_id, prob.quad)
    sym_first_ms = 1e3 * (time.perf_counter() - t0)
    params = prob.form.params()
    alg = algorithmic_bytes(prob, nnz)
    small = alg < 2 * L2_BYTES
    sampler = ClockSampler(local)
    sampler.start()
    ms = time_numeric(torch, eng, stream, params, steps, warmup, small) / steps
    clocks = sampler.result()
    nel = prob.meshes[0].nel_
    out = {"workload": f"{WORKLOADS[workload][1]}, N={n}", "elements": nel, "ndofs": int(prob.ndofs), "nnz": int(nnz),
           "ms_per_step": ms, "value": nel / (ms / 1e3), "unit": "elements/s",
           "roofline": {"achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "frac": alg / (ms / 1e3) / 1e9 / peak, "unit": "GB/s",
                        "algorithmic_bytes_per_element": alg / nel},
           "halo_factor": eng.stat(_lib.STAT_TILE_ELEMS) / max(nel, 1), "tiles": int(eng.stat(_lib.STAT_NTILES)),
           "tile_elems": -(-nel // max(int(eng.stat(_lib.STAT_NTILES)), 1)),
           "symbolic_first_call_ms": sym_first_ms, "mesh_generation_s": gen_s,
           "l2": "working set below 2 x L2: a 252 MB buffer is written between timed launches" if small else "inputs larger than L2",
           "clocks": clocks, "device_bytes": eng.stat(_lib.STAT_DEVICE_BYTES)}
    eng.close()
    del meshes, dofs, prob
    torch.cuda.empty_cache()
    return out


def _as_torch(torch, dev_array):
    return torch.as_tensor(dev_array, device="cuda")


def config5_checksums(torch, eng, band, chunk_cols=1 << 24):
    """(nnz, pattern checksum mod 2^64, sum of nzval^2) of this rank's device-resident CSC block, computed with torch ops on
    zero-copy views (efg_device_csc); columns are walked in chunks so the temporaries stay below ~2 GB."""
    from elfel_jl_b200.sharding import _mix
    cp, rv, nz = (_as_torch(torch, a) for a in eng.device_csc())
    cols = torch.cat([torch.arange(int(f), int(l) + 1, dtype=torch.int64, device="cuda") for f, l in zip(band.firsts, band.lasts)])
    ncl = cols.numel()
    chk, sq = 0, 0.0
    for a in range(0, ncl, chunk_cols):
        b = min(a + chunk_cols, ncl)
        ptr = cp[a:b + 1]
        lo, hi = int(ptr[0].item()) - 1, int(ptr[-1].item()) - 1
        cnt = ptr[1:] - ptr[:-1]
        col_of = torch.repeat_interleave(cols[a:b], cnt)
        rows = rv[lo:hi].to(torch.int64) + 1
        chk = (chk + int(_mix(rows, col_of).sum().item())) % (1 << 64)
        v = nz[lo:hi]
        sq += float((v * v).sum().item())
        del col_of, rows, cnt, v
    return int(cp[-1].item()) - 1, chk, sq


def run_config5(torch, dist, efg, _lib, args, rank, world, local, peak):
    """BASELINE config 5, strong scaling: the 16384 x 16384 Q4 heat problem split into `world` horizontal bands (owner
    computes, one halo cell row per cut, no data-path collective).  Every rank generates ONLY its band on its GPU
    (closed-form numbering, sharding.block_band) and hands device pointers to the library."""
    from elfel_jl_b200 import sharding as sh
    N = args.config5_n
    dev = torch.device("cuda", local)
    rec = {"workload": f"heat Poisson FEH1_Q4 Gauss order 2, {N} x {N} Q4block (BASELINE config 5), columns split over {world} rank(s)",
           "scaling": "strong", "n_gpus": world, "elements": N * N, "nnz_expected": (3 * N + 1) ** 2}

    def assemble_band(r, w, steps, warmup, validate):
        t0 = time.perf_counter()
        band = sh.block_band(efg.Q4, N, r, w, dev)
        torch.cuda.synchronize()
        gen_s = time.perf_counter() - t0
        eng = efg.Engine(local)
        stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
        eng.set_mesh(0, efg.Q4, band.conn, band.xy)
        eng.set_space(0, 0, band.dofnums)
        eng.start(band.ndofs, band.ndofs)
        eng.set_column_ranges(band.firsts, band.lasts)
        eng.synchronize()
        t0 = time.perf_counter()
        nnz = eng.symbolic(_lib.FORM_HEAT, 2)
        sym_ms = 1e3 * (time.perf_counter() - t0)
        nel, nnodes = int(band.conn.shape[0]), int(band.xy.shape[0])
        ms = time_numeric(torch, eng, stream, [1.0], steps, warmup, False) / steps
        alg = 4 * 4 * nel + 16 * nnodes + 4 * nnodes + 4 * 16 * nel + 8 * nnz
        out = {"rank_elements_incl_shard_halo": nel, "owned_cell_rows": band.rows[1] - band.rows[0] - (1 if band.rows[1] == N + 1 else 0),
               "nnz": int(nnz), "numeric_ms": ms, "symbolic_first_call_ms": sym_ms, "mesh_generation_s": gen_s,
               "tile_halo_factor": eng.stat(_lib.STAT_TILE_ELEMS) / nel, "tiles": int(eng.stat(_lib.STAT_NTILES)),
               "device_bytes": eng.stat(_lib.STAT_DEVICE_BYTES), "algorithmic_bytes": int(alg),
               "achieved_GBps": alg / (ms / 1e3) / 1e9, "frac": alg / (ms / 1e3) / 1e9 / peak}
        val = None
        if validate:
            got = config5_checksums(torch, eng, band)
            exp = sh.q4_expected_checksums(N, band.rows[0], band.rows[1], dev)
            val = (got, exp)
        eng.close()
        del band
        torch.cuda.empty_cache()
        return out, val

    if world > 1:
        dist.barrier()
    mine, (got, exp) = assemble_band(rank, world, args.steps, args.warmup, True)
    ms = mine["numeric_ms"]
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        # validation over NCCL, outside the timed region: nnz and sum of squares (float64 sums), the pattern checksum as two
        # 32-bit halves (exact in int64), per-rank nnz gathered -> offsets of the rank blocks in the concatenated CSC
        red = torch.tensor([float(got[0]), got[2], float(exp[0]), exp[2], float(mine["rank_elements_incl_shard_halo"])], device="cuda", dtype=torch.float64)
        dist.all_reduce(red)
        halves = torch.tensor([got[1] & 0xFFFFFFFF, got[1] >> 32, exp[1] & 0xFFFFFFFF, exp[1] >> 32], device="cuda", dtype=torch.int64)
        dist.all_reduce(halves)
        per_rank = [torch.zeros(3, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(per_rank, torch.tensor([float(got[0]), ms, mine["tile_halo_factor"]], device="cuda", dtype=torch.float64))
        nnz_g, sq_g, nnz_e, sq_e, nel_sum = (float(x) for x in red.tolist())
        h = [int(x) for x in halves.tolist()]
        chk_g, chk_e = (h[0] + (h[1] << 32)) % (1 << 64), (h[2] + (h[3] << 32)) % (1 << 64)
        rank_nnz = [int(x[0].item()) for x in per_rank]
        rank_ms = [float(x[1].item()) for x in per_rank]
    else:
        ms_max, nnz_g, sq_g, nnz_e, sq_e, nel_sum = ms, float(got[0]), got[2], float(exp[0]), exp[2], float(mine["rank_elements_incl_shard_halo"])
        chk_g, chk_e, rank_nnz, rank_ms = got[1], exp[1], [got[0]], [ms]
    offs = np.concatenate([[0], np.cumsum(rank_nnz)]).astype(np.int64)
    rec.update({
        "numeric_ms": ms_max, "value": N * N / (ms_max / 1e3), "unit": "elements/s", "steps": args.steps, "warmup": args.warmup,
        "rank_numeric_ms": rank_ms, "rank_nnz": rank_nnz, "rank_block_offsets": [int(x) for x in offs[:-1]],
        "shard_halo_fraction": nel_sum / (N * N) - 1.0, "rank0": mine,
        "validation": {"nnz": int(nnz_g), "nnz_exact": int(nnz_g) == (3 * N + 1) ** 2 and int(nnz_e) == (3 * N + 1) ** 2,
                       "pattern_checksum": f"{chk_g:016x}", "pattern_checksum_expected": f"{chk_e:016x}", "pattern_checksum_ok": chk_g == chk_e,
                       "sum_nzval_sq": sq_g, "sum_nzval_sq_expected": sq_e, "sum_nzval_sq_rel_err": abs(sq_g - sq_e) / sq_e,
                       "how": "per rank on its device block; expected values derived from the grid alone (9-point node adjacency, uniform "
                              "square cells); reduced with NCCL all_reduce / all_gather outside the timed region"},
        "roofline": {"frac_of_measured_peak_per_gpu": (4 * 4 + 16 * (N + 1) ** 2 / N ** 2 + 4 * (N + 1) ** 2 / N ** 2 + 64 + 8 * (3 * N + 1) ** 2 / N ** 2)
                     * N * N / world / (ms_max / 1e3) / 1e9 / peak, "algorithmic_bytes_per_element": 172.0}})
    if world == 1 and not args.no_others:
        # what one rank of a 2/4/8-way split does, timed on this GPU: parallel efficiency = whole / (n * share)
        shares = {}
        for w in (2, 4, 8):
            o, _ = assemble_band(w // 2, w, max(5, args.steps // 2), 3, False)
            shares[str(w)] = {"numeric_ms": o["numeric_ms"], "rank_elements_incl_shard_halo": o["rank_elements_incl_shard_halo"],
                              "predicted_efficiency": ms_max / (w * o["numeric_ms"])}
        rec["one_gpu_shares"] = shares
    return rec


def run_verify(args):
    """--verify: BASELINE configs 2, 3 and 4 at FULL size on the GPU, checked against the direct-accumulate oracle on a spread
    of column blocks (bit-exact pattern; nzval within 1e-12 relative / 1e-14 absolute in the default FP mode and == in the
    strict mode).  One JSON line with the largest errors seen."""
    import torch
    import elfel_jl_b200 as efg
    from elfel_jl_b200 import _lib
    from oracle import oracle as orc
    from concurrent.futures import ThreadPoolExecutor
    orc.build()
    res = []
    for wl, n in (("heat_t6", 4000), ("elasticity_t6", 2000), ("stokes_gen", 1000)):
        if args.workload != "heat_t6" and wl != args.workload:
            continue
        n = args.n or n
        t0 = time.perf_counter()
        prob = make_problem(efg, wl, n)
        nd = int(prob.ndofs)
        pairs = [(prob.meshes[ms].conn, sp.field.dofnums) for sp, ms in zip(prob.spaces, prob.space_mesh)]
        dconn = [(torch.from_numpy(np.ascontiguousarray(c)).cuda(), torch.from_numpy(np.ascontiguousarray(d)).cuda()) for c, d in pairs]
        gen_s = time.perf_counter() - t0
        # column blocks: the first and last free columns, the data-dof tail, and blocks spread over the rest
        width = max(1000, min(args.verify_cols, nd))
        starts = sorted(set([1, max(1, nd - width + 1)] + [max(1, int(x)) for x in np.linspace(1, max(1, nd - width + 1), args.verify_blocks)]))
        blocks = [(a, min(a + width - 1, nd)) for a in starts]
        eng = efg.Engine(0)
        efg.load_problem(eng, prob)
        rec = {"workload": f"{WORKLOADS[wl][1]}, N={n}", "elements": int(prob.nel), "ndofs": nd, "column_blocks": len(blocks), "columns_checked": 0,
               "nonzeros_checked": 0, "pattern_bit_exact": True, "default_fp": {"max_abs": 0.0, "max_rel": 0.0, "within_1e-12_1e-14": True},
               "strict_fp": {"max_abs": 0.0, "identical": True}}
        gpu = {}
        for strict in (0, 1):
            eng.set_option(_lib.OPT_STRICT_FP, strict)
            nnz = eng.assemble(prob.form.form_id, prob.quad, prob.form.params())
            gpu[strict] = [eng.block(1, nd, a, b) for a, b in blocks]
        rec["nnz"] = int(nnz)

        def elems_of(a, b):
            hit = None
            for c, d in dconn:
                inb = ((d >= a) & (d <= b)).any(dim=1)
                h = inb[c - 1].any(dim=1)
                hit = h if hit is None else (hit | h)
            return torch.nonzero(hit).reshape(-1).cpu().numpy().astype(np.int64)
        elists = [elems_of(a, b) for a, b in blocks]

        def oracle_block(k):
            a, b = blocks[k]
            return orc.assemble_direct(*efg.oracle_args(prob), nd, nd, c0=a, c1=b, elist=elists[k])
        t0 = time.perf_counter()
        with ThreadPoolExecutor(min(os.cpu_count() or 1, 16)) as ex:
            want = list(ex.map(oracle_block, range(len(blocks))))
        rec["oracle_s"] = time.perf_counter() - t0
        for k, (ocp, orv, onz) in enumerate(want):
            for strict in (0, 1):
                cp, rv, nz = gpu[strict][k]
                ok = np.array_equal(cp, ocp) and np.array_equal(rv, orv)
                rec["pattern_bit_exact"] = rec["pattern_bit_exact"] and bool(ok)
                if not ok:
                    continue
                d = np.abs(nz - onz)
                if strict:
                    rec["strict_fp"]["max_abs"] = max(rec["strict_fp"]["max_abs"], float(d.max(initial=0.0)))
                    rec["strict_fp"]["identical"] = rec["strict_fp"]["identical"] and bool(np.array_equal(nz, onz))
                else:
                    rel = d / np.maximum(np.abs(onz), 1e-300)
                    rec["default_fp"]["max_abs"] = max(rec["default_fp"]["max_abs"], float(d.max(initial=0.0)))
                    rec["default_fp"]["max_rel"] = max(rec["default_fp"]["max_rel"], float(rel[np.abs(onz) > 1e-14].max(initial=0.0)))
                    rec["default_fp"]["within_1e-12_1e-14"] = rec["default_fp"]["within_1e-12_1e-14"] and bool(np.all(d <= 1e-14 + 1e-12 * np.abs(onz)))
            rec["columns_checked"] += blocks[k][1] - blocks[k][0] + 1
            rec["nonzeros_checked"] += int(len(orv))
        rec["mesh_generation_s"] = gen_s
        eng.close()
        del dconn, gpu, want
        torch.cuda.empty_cache()
        res.append(rec)
    print(json.dumps({"verify": res, "oracle": "oracle/elfel_oracle.c direct-accumulate mode (== the COO + sparse() restatement, tests/test_oracle_golden.py)",
                      "tolerance": "pattern bit-exact; nzval 1e-12 relative / 1e-14 absolute (default FP mode), == (EFG_OPT_STRICT_FP)"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="heat_t6", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="mesh subdivisions per side (0 = the BASELINE size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-n", type=int, default=1000, help="mesh size of the CPU arm (--impl reference and cpu_baseline use the same N)")
    ap.add_argument("--cpu-n", type=int, default=0, help="0 = --ref-n")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-warmup", type=int, default=4, help="untimed e2e calls (the device memory pool needs a few calls to reach steady state)")
    ap.add_argument("--tile-elems", type=int, default=0)
    ap.add_argument("--sfc", type=int, default=1)
    ap.add_argument("--strict", type=int, default=0)
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--host-widen", type=int, default=-1, help="e2e: 1 = Int32 row indices widened by host threads, 0 = widened on the device, -1 = 1 for <= 2 ranks")
    ap.add_argument("--no-defer-xy", action="store_true", help="e2e: copy the coordinates inside efg_set_mesh instead of overlapping them with the pattern kernels")
    ap.add_argument("--no-callers", action="store_true", help="skip the f1/f2 rows (load vector, K*x) timed after the hot path")
    ap.add_argument("--no-config5", action="store_true", help="skip the config 5 strong-scaling record")
    ap.add_argument("--config5-n", type=int, default=16384)
    ap.add_argument("--no-others", action="store_true", help="skip other_configs (N = 1 only)")
    ap.add_argument("--no-widened", action="store_true", help="skip widened_rows (SURVEY 8f rows f4 / f5 timed like the other configs)")
    ap.add_argument("--only-config5", action="store_true", help="print only the config 5 record (development)")
    ap.add_argument("--verify", action="store_true", help="full-size parity of configs 2/3/4 against the oracle on column blocks (one JSON line)")
    ap.add_argument("--verify-blocks", type=int, default=10)
    ap.add_argument("--verify-cols", type=int, default=40000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    args.cpu_n = args.cpu_n or args.ref_n
    if args.impl == "reference":
        return run_reference(args)
    if args.verify:
        return run_verify(args)

    import torch
    import elfel_jl_b200 as efg
    from elfel_jl_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the assembly path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        # (a rank that fails must not leave the others waiting in a collective until the job limit)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=300))
    peak, peak_src = load_peak()

    if args.only_config5:
        rec = run_config5(torch, dist, efg, _lib, args, rank, world, local, peak)
        if rank == 0:
            print(json.dumps({"config5": rec}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    n = args.n or WORKLOADS[args.workload][0]
    if world > 1:
        # weak scaling: a mesh `world` units tall, rank r owns the dofs of horizontal band r (a set of column
        # ranges), halo elements replicated; generated on the GPU (elfel.jl_b200/sharding.py)
        from elfel_jl_b200.sharding import shard_problem
        prob, col_range, nel_global = shard_problem(efg, args.workload, n, rank, world)
    else:
        prob, col_range, nel_global = make_problem(efg, args.workload, n), None, None
    nel_global = nel_global or prob.nel

    torch.cuda.synchronize()
    t_create = time.perf_counter()
    eng = efg.Engine(local)
    create_ms = 1e3 * (time.perf_counter() - t_create)       # efg_create: stream, events, the library's device module
    eng.set_option(_lib.OPT_PATH, args.path)
    eng.set_option(_lib.OPT_STRICT_FP, args.strict)
    eng.set_option(_lib.OPT_TILE_ELEMS, args.tile_elems)
    eng.set_option(_lib.OPT_SFC_ORDER, args.sfc)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
    params = prob.form.params()

    # host buffers (pinned) of the reference-facing call
    def pin(a):
        t = a.detach().cpu().contiguous() if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        return t.pin_memory()
    h_mesh = [(m.kind, pin(m.conn), pin(m.xy)) for m in prob.meshes]
    h_dofs = [pin(s.field.dofnums) for s in prob.spaces]
    for m in prob.meshes:           # the host copies are the inputs from here on
        m.conn_shape, m.nnodes_, m.nel_ = tuple(m.conn.shape), int(m.xy.shape[0]), int(m.conn.shape[0])
    if world > 1:
        for m in prob.meshes:
            m.conn = m.xy = None
        for sp in prob.spaces:
            sp.field.dofnums = None
        torch.cuda.empty_cache()
    h2d = sum(c.numel() * 8 + x.numel() * 8 for _, c, x in h_mesh) + sum(d.numel() * 8 for d in h_dofs)
    prob.ndofs_local_ = int(sum(d.numel() for d in h_dofs))

    def load():
        for slot, (kind, c, x) in enumerate(h_mesh):
            eng.set_mesh(slot, kind, c, x)
        for slot, (d, ms) in enumerate(zip(h_dofs, prob.space_mesh)):
            eng.set_space(slot, ms, d)
        eng.start(prob.ndofs, prob.ndofs)
        if col_range is not None:
            eng.set_column_ranges(*col_range)

    def rank_max(dt):
        if world > 1:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return dt

    # ---- e2e through the C ABI with host buffers.  The FIRST call of the process is timed too (`e2e_first_call`: what a
    # script that assembles once sees; the pinned output arrays are allocated between its two timed parts, after nnz is
    # known, like the caller of the two-call pattern does) --------------------------------------------------------------
    e2e = None
    fid, quad = prob.form.form_id, prob.quad
    # row indices: Int32 over PCIe + host-side widening while one or two ranks share the host, device-side widening beyond
    # (8 ranks: the host's memory system is the bottleneck; measured 1296 ms with host widening at 8 GPUs)
    host_widen = args.host_widen if args.host_widen >= 0 else int(world <= 2)
    eng.set_option(_lib.OPT_HOST_WIDEN, host_widen)
    if not args.no_defer_xy:
        eng.set_option(_lib.OPT_DEFER_XY, 1)     # the pinned inputs stay alive across the sequence (like the Julia shim's GC.@preserve)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    load()
    nnz = eng.pattern(fid, quad)
    first_a = time.perf_counter() - t0
    if args.no_e2e:
        eng.assemble(fid, quad, params)
        eng.synchronize()
        sym_ms = eng.stat(_lib.STAT_SYMBOLIC_MS)
    else:
        ncl = eng.ncols_local
        # (allocated pinned directly: .pin_memory() on a pageable tensor would allocate, copy and free another 12 GB here)
        o_colptr = torch.empty(ncl + 1, dtype=torch.int64, pin_memory=True)
        o_rowval = torch.empty(nnz, dtype=torch.int64, pin_memory=True)
        o_nzval = torch.empty(nnz, dtype=torch.float64, pin_memory=True)
        d2h = 8 * (ncl + 1) + (4 if host_widen else 8) * nnz + 8 * nnz      # colptr Int64, rowval (Int32 when widened by host threads), nzval
        t0 = time.perf_counter()
        eng.fetch_pattern_async(o_colptr, o_rowval)
        eng.numeric(params)
        eng.fetch_csc(None, None, o_nzval)
        eng.synchronize()
        first_s = rank_max(first_a + time.perf_counter() - t0)
        sym_ms = eng.stat(_lib.STAT_SYMBOLIC_MS)
        ts, all_ts, sym_hist = [], [], []
        for it in range(args.e2e_warmup + args.e2e_steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            load()
            eng.pattern(fid, quad)                          # -> nnz: the caller sizes its arrays here
            eng.fetch_pattern_async(o_colptr, o_rowval)     # structure travels while the tiles and the values are computed
            eng.numeric(params)
            eng.fetch_csc(None, None, o_nzval)
            eng.synchronize()
            dt = rank_max(time.perf_counter() - t0)
            all_ts.append(round(1e3 * dt, 1))
            sym_hist.append(round(eng.stat(_lib.STAT_SYMBOLIC_MS), 1))
            if it >= args.e2e_warmup:
                ts.append(dt)
        # the plain sequence (everything in one call, then one fetch) for comparison
        ps = []
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            load()
            eng.assemble(fid, quad, params)
            eng.fetch_csc(o_colptr, o_rowval, o_nzval)
            eng.synchronize()
            ps.append(rank_max(time.perf_counter() - t0))
        e2e_s = float(np.mean(ts))
        # re-assembly on the cached pattern (time stepping / Newton): efg_numeric + the values only
        rs = []
        for it in range(1 + args.e2e_steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.numeric(params)
            eng.fetch_csc(None, None, o_nzval)
            eng.synchronize()
            dt = rank_max(time.perf_counter() - t0)
            if it >= 1:
                rs.append(dt)
        re_s = float(np.mean(rs))
        e2e = {"value": nel_global / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s, "steps": args.e2e_steps, "warmup": args.e2e_warmup,
               "all_calls_ms": all_ts, "symbolic_ms_per_call": sym_hist,
               "row_index_widening": "host threads, Int32 over PCIe (EFG_OPT_HOST_WIDEN = 1)" if host_widen else "device, Int64 over PCIe (EFG_OPT_HOST_WIDEN = 0: many ranks share the host)",
               "what": "efg_set_mesh/_space (EFG_OPT_DEFER_XY: coordinates copied while the pattern kernels run) + efg_start + efg_pattern + efg_fetch_pattern_async + efg_numeric + efg_fetch_csc(nzval), pinned host buffers "
                       "(the sequence of the Julia shim's assemble!/finish!: structure copied out while tiles and values are computed)",
               "plain_sequence_ms": 1e3 * float(np.min(ps)),
               "e2e_first_call": {"value": nel_global / first_s, "ms": 1e3 * first_s, "symbolic_ms": sym_ms, "efg_create_ms": create_ms,
                                  "what": "the same sequence as the first library call of the process (no warm-up; CUDA context creation excluded; efg_create, which loads the library's device module, is outside and reported as efg_create_ms)"},
               "e2e_reassembly": {"value": nel_global / re_s, "ms": 1e3 * re_s, "d2h_bytes_per_step": int(8 * nnz),
                                  "what": "efg_numeric on the cached pattern + efg_fetch_csc(nzval only)"}}
        checksum = float(o_nzval.sum().item())
        del o_colptr, o_rowval, o_nzval
    if args.no_e2e:
        checksum = None

    # ---- device-resident numeric phase (the hot path), CUDA events on the library's stream --------------
    alg = algorithmic_bytes(prob, nnz)     # this rank's launch (its sub-mesh incl. replicated halo elements)
    small = alg < 2 * L2_BYTES
    for _ in range(args.warmup):
        eng.numeric(params)
    eng.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches1 = eng.stat(_lib.STAT_KERNEL_LAUNCHES)
    ms_total = time_numeric(torch, eng, stream, params, args.steps, 0, small)
    torch.cuda.synchronize()
    clocks = sampler.result()
    gpu_launches = int(eng.stat(_lib.STAT_KERNEL_LAUNCHES) - launches1)
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        tn = torch.tensor([float(nnz)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tn)
        nnz_global = int(tn.item())
    else:
        nnz_global = nnz
    ms_step = ms_total / args.steps
    value = nel_global / (ms_step / 1e3)

    path = int(eng.stat(_lib.STAT_PATH))
    # per-launch time of this rank's kernel: the timed region holds only numeric kernels
    achieved = alg / (ms_step / 1e3) / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        traffic = json.load(open(tf)).get(f"{args.workload}_N{n}")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg),
                "algorithmic_bytes_per_element": alg / prob.meshes[0].nel_,
                "designed_bytes_per_launch": int(eng.stat(_lib.STAT_NUMERIC_BYTES)),
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "kernel": "k_tl_numeric" if path == 2 else "k_tp_elem_matrices+k_tp_gather"}

    # ---- SURVEY 8f rows f1 / f2, timed beside the hot path (heat workloads, one GPU): the load vector of the same
    # integrate! loop and K*T of the examples' solve!, device-resident, CUDA events inside the library ----------
    callers = None
    st_ntiles, st_tile_elems = int(eng.stat(_lib.STAT_NTILES)), eng.stat(_lib.STAT_TILE_ELEMS)
    st_dev_bytes = eng.stat(_lib.STAT_DEVICE_BYTES)
    if world == 1 and prob.form.form_id == 1 and not args.no_callers and path == 2:
        m0 = prob.meshes[0]
        nd = int(prob.ndofs)
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [-6.0], nd)          # builds the dof -> contribution map
        tv = []
        for _ in range(5):
            eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [-6.0], nd)
            tv.append(eng.stat(_lib.STAT_VEC_MS))
        v_ms = float(np.median(tv))
        v_alg = 4 * m0.kind * m0.nel_ + 16 * m0.nnodes_ + 4 * nd + 4 * m0.kind * m0.nel_ + 8 * nd
        xd = torch.ones(nd, dtype=torch.float64, device="cuda")
        yd = torch.empty(nd, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        eng.spmv(xd, yd)                                                       # builds the row-major view
        ts = []
        for _ in range(5):
            eng.spmv(xd, yd)
            ts.append(eng.stat(_lib.STAT_SPMV_MS))
        s_ms = float(np.median(ts))
        s_alg = 12 * nnz + 8 * nd + 8 * nd + 8 * (nd + 1)
        callers = {
            "load_vector": {"what": "efg_vec_assemble (SysvecAssembler: fe[j] += N[j]*Q*JxW), 2 kernels, dof map cached",
                            "ms": v_ms, "elements_per_s": m0.nel_ / (v_ms / 1e3), "algorithmic_bytes": int(v_alg),
                            "achieved_GBps": v_alg / (v_ms / 1e3) / 1e9, "frac": v_alg / (v_ms / 1e3) / 1e9 / peak},
            "spmv": {"what": "efg_spmv (KT = K*T, SparseArrays summation order); heat matrices are bitwise symmetric: CSC columns read in place as rows",
                     "ms": s_ms, "nnz_per_s": nnz / (s_ms / 1e3), "algorithmic_bytes": int(s_alg),
                     "achieved_GBps": s_alg / (s_ms / 1e3) / 1e9, "frac": s_alg / (s_ms / 1e3) / 1e9 / peak,
                     "y_checksum": float(yd.sum().item())},
        }
        del xd, yd
        # K and F from ONE pass (efg_numeric_with_load): a second tiling with room for the element load vector in the stage
        try:
            eng.set_option(_lib.OPT_FUSE_LOAD, 1)
            eng.symbolic(fid, quad)
            for _ in range(3):
                eng.numeric_with_load(params, -6.0)
            tf = []
            for _ in range(5):
                eng.numeric_with_load(params, -6.0)
                eng.synchronize()
                tf.append(eng.stat(_lib.STAT_NUMERIC_MS))
            f_ms = float(np.median(tf))
            callers["fused_matrix_and_load_vector"] = {
                "what": "efg_numeric_with_load: K and F of one integrate! pass in ONE kernel (fe staged next to ke, gathered through the "
                        "diagonal nonzeros' contribution lists); compare with numeric + load_vector",
                "ms": f_ms, "separate_ms": ms_step + v_ms, "extra_ms_over_matrix_only": f_ms - ms_step,
                "tile_elems": -(-int(m0.nel_) // max(int(eng.stat(_lib.STAT_NTILES)), 1))}
        except Exception as e:          # noqa: BLE001 -- a record, not the headline
            callers["fused_matrix_and_load_vector"] = {"error": f"{type(e).__name__}: {e}"}

    out = {
        "metric": "elements assembled/s", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOADS[args.workload][1]}, N={n}" + (f", columns split over {world} ranks" if world > 1 else ""),
                   "elements": int(nel_global), "ndofs": int(prob.ndofs), "nnz": int(nnz_global),
                   "l2": "working set below 2 x L2: a 252 MB buffer is written between timed launches (each launch has its own event pair)"
                         if small else "inputs larger than L2",
                   "path": {1: "two-pass", 2: "tiled-fused"}[path], "strict_fp": args.strict,
                   "tile_elems": int(args.tile_elems) or -(-int(prob.meshes[0].nel_) // max(st_ntiles, 1)),
                   "tile_elems_source": "option" if args.tile_elems else "automatic (largest size with two CTAs per SM)", "sfc_order": args.sfc,
                   "tiles": st_ntiles,
                   "halo_factor": st_tile_elems / max(prob.meshes[0].nel_, 1),
                   "rank_elements_incl_shard_halo": int(prob.meshes[0].nel_)},
        "roofline": roofline, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
        "phases": {"symbolic_ms": sym_ms, "symbolic_ms_is": "first call of the process", "numeric_ms": ms_step,
                   "value_with_symbolic": nel_global / ((ms_step + sym_ms) / 1e3)},
        "device_bytes": st_dev_bytes, "nzval_checksum": checksum, "next_rows": callers,
    }
    eng.close()
    del h_mesh, h_dofs, prob
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, timed the same way (N = 1 runs) ------------------------------------------------
    if world == 1 and not args.no_others and args.workload == "heat_t6" and not args.n:
        others = []
        for wl, nn in (("elasticity_t6", 2000), ("stokes_gen", 1000), ("heat_q4", 5792), ("heat_t3", 100)):
            try:
                others.append(bench_device_problem(torch, efg, _lib, local, wl, nn, args.steps, args.warmup, peak))
            except Exception as e:      # keep the headline line even if a side record fails
                others.append({"workload": wl, "error": f"{type(e).__name__}: {e}"})
        out["other_configs"] = others
        if not args.no_widened:
            out["widened_rows"] = bench_widened_rows(torch, efg, _lib, local, args.steps, args.warmup, peak)
    if not args.no_config5 and args.workload == "heat_t6" and not args.n:
        try:
            out["config5"] = run_config5(torch, dist, efg, _lib, args, rank, world, local, peak)
        except Exception as e:
            if world > 1:
                raise           # a rank that drops out of the collectives would hang the others: fail loudly instead
            out["config5"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(efg, args.workload, min(args.cpu_n, n))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
