#!/usr/bin/env python
"""bench.py — NCMC steps/s on the T4 lysozyme L99A – toluene workload (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--replicas R] [--impl reference]

A "step" is one pass of the hot path — one NCMC integrator step (H V R O R V H with work accumulation) of every
walker held by the GPU.  N > 1 is launched by torchrun, one rank per GPU; walkers are independent (weak scaling,
no data-path collective), only work/acceptance statistics are gathered over NCCL after the timed region.

Printed JSON (one line, rank 0): value = walker-steps/s with the state resident in HBM, CUDA-event timed, max over
ranks; e2e = the same metric through the public Context API with host buffers (H2D of positions + velocities, K
steps with the on-device rotation move, D2H of positions + protocol work, Metropolis test); roofline for the
dominant kernel from live per-kernel CUDA-event timing (direct-launch profiling pass); cpu_baseline = the
oracle's C twin (reference semantics, all host cores) on a bounded sample.  `--impl reference` times that CPU
implementation alone (the reference stack — OpenMM/openmmtools/parmed — is not installable here).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 60
FP32_PEAK_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
FUNCS = {'lambda_sterics': 'min(1, (1/0.3)*abs(lambda-0.5))',
         'lambda_electrostatics': 'step(0.2-lambda) - 1/0.2*lambda*step(0.2-lambda) + 1/0.2*(lambda-0.8)*step(lambda-0.8)'}

# SURVEY.md §8(d) measurement configurations.  `t4l` (M2 = BASELINE configs[1]) is the bench line the driver reads;
# the others are the same engine on the other configurations, selected with --workload.
WORKLOADS = {
    't4l': dict(
        case='t4l_surrogate', nsteps_nc=5000, dt=0.004, move='rotate', replicas=1,
        p_in=4672867,         # non-excluded pairs within 1.0 nm at the fixture coordinates (oracle count)
        text='T4L-toluene geometry (22340 atoms, surrogate force field: eqToluene.prmtop is missing upstream), explicit '
             'TIP3P, PME rc 1.0 nm tol 5e-3 grid 24x25x28, HBonds + rigid water, HMR 3.024 Da, dt 4 fs, 300 K, '
             'nstepsNC=5000, RandomLigandRotationMove at moveStep'),
    'tolparm': dict(
        case='tol_parm', nsteps_nc=100, dt=0.002, move='rotate', replicas=1, p_in=None,
        text='M1 / BASELINE configs[0]: toluene in TIP3P (TOL-parm.prmtop, 975 atoms, cubic 2.1786 nm), PME rc 0.8 nm '
             'tol 5e-4 grid 24^3, HBonds, dt 2 fs, 300 K, nstepsNC=100, RandomLigandRotationMove at moveStep'),
    'water': dict(
        case='t4l_surrogate', nsteps_nc=1000, dt=0.002, move='water', replicas=1, p_in=4672867,
        alch=[2657, 2658, 2659], selection='(index 1656) or (index 1657)', radius_nm=0.9,
        text='M4 / BASELINE configs[3]: WaterTranslationMove on the T4L geometry (22340 atoms, surrogate force field), '
             'alchemical water = first HOH (atoms 2657-2659), sphere 0.9 nm around atoms 1656/1657, nstepsNC=1000, '
             'dt 2 fs, swap / translate / check hooks on the device'),
    'm5': dict(
        case='tol_parm', tile=(6, 6, 7), nsteps_nc=5000, dt=0.002, move='rotate', replicas=8, p_in=None,
        kw=dict(cutoff_angstrom=10.0, ewaldErrorTolerance=0.005),
        text='M5 / BASELINE configs[4]: TOL-parm tiled 6x6x7 = 245700 atoms, box 13.07x13.07x15.25 nm, PME rc 1.0 nm '
             'tol 5e-3, HBonds, dt 2 fs, 300 K, nstepsNC=5000, one alchemical toluene, 8 walkers per GPU'),
}


def load_workload(name='t4l'):
    """Structure, alchemical System, flat topology and start coordinates (nm) of a measurement configuration."""
    from tests.gpu_checks import CASES, GOLDEN, tile_structure
    from blues_b200 import unit as u
    from blues_b200.structure import Structure
    from blues_b200.alchemy import AbsoluteAlchemicalFactory, AlchemicalRegion
    w = WORKLOADS[name]
    base = Structure.load_npz(os.path.join(GOLDEN, w['case'] + '.npz'))
    s = tile_structure(base, w['tile']) if w.get('tile') else base
    kw = dict(CASES[w['case']]['kw'])
    over = dict(w.get('kw', {}))
    if 'cutoff_angstrom' in over:
        kw['nonbondedCutoff'] = over.pop('cutoff_angstrom') * u.angstroms
    kw.update(over)
    system = s.createSystem(**kw)
    alch = w.get('alch', CASES[w['case']]['alch'])
    system = AbsoluteAlchemicalFactory().create_alchemical_system(system, AlchemicalRegion(alchemical_atoms=alch))
    return dict(w, name=name, structure=s, base_structure=base, system=system, topo=system.flatten(),
                x=s.coordinates * 0.1, alch=alch)


def lambda_tables_for_cpu(nsteps_nc):
    """lambda tables for the CPU leg (the native arm evaluates the same expressions inside the integrator object)"""
    from tests.gpu_checks import lambda_tables
    return lambda_tables(nsteps_nc)


def make_move(wl):
    """The move object of the workload and its on-device descriptor."""
    from blues_b200 import unit
    from blues_b200.moves import RandomLigandRotationMove, WaterTranslationMove
    if wl['move'] == 'water':
        mv = WaterTranslationMove(wl['structure'], protein_selection=wl['selection'], radius=wl['radius_nm'] * unit.nanometers)
        mv.atom_indices = list(wl['alch'])
        return mv
    return RandomLigandRotationMove(wl['base_structure'], 'LIG')     # tiled boxes: the first copy's toluene


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted(set(n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith('active')))
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


def cpu_relax(topo, x, iters=60, max_disp=0.005):
    """Capped steepest descent with the constraints re-imposed (CPU oracle): the TOL-parm start coordinates carry
    the type-index quirk documented in DESIGN.md §8 and, like the reference's tests, need minimising before dynamics."""
    from oracle.c_oracle import COracle
    from oracle.ncmc_oracle import Constraints
    c = COracle(topo)
    cons = Constraints(topo)
    mobile = (np.asarray(topo['mass']) > 0)[:, None]
    x = cons.apply_positions(np.asarray(x, float), np.asarray(x, float), tol=1e-10)
    e, f = c.energy_forces(x)[:2]
    step = 1e-5
    for _ in range(iters):
        d = step * f * mobile
        n = np.linalg.norm(d, axis=1, keepdims=True)
        d *= np.minimum(1.0, max_disp / np.maximum(n, 1e-30))
        xn = cons.apply_positions(x + d, x, tol=1e-10)
        en, fn = c.energy_forces(xn)[:2]
        if np.isfinite(en) and en < e:
            x, e, f, step = xn, en, fn, step * 1.3
        else:
            step *= 0.4
    return x


def cpu_reference_run(wl, steps, warmup, budget_s, x=None):
    """The reference path on the host: oracle C twin, reference semantics (3 evaluations/step), all cores."""
    from oracle.c_oracle import COracle
    ls, le = lambda_tables_for_cpu(wl['nsteps_nc'])
    c = COracle(wl['topo'], ls, le, 'H V R O R V H', 300.0, 1.0, wl['dt'], wl['nsteps_nc'], 1, 0.2, 0.8, seed=20261017)
    if x is None and wl['case'] == 'tol_parm':
        # relax one periodic image on the CPU, then tile the relaxed coordinates
        from tests.gpu_checks import CASES
        base = wl['base_structure']
        kw = dict(CASES['tol_parm']['kw'])
        xb = cpu_relax(base.createSystem(**kw).flatten(), base.coordinates * 0.1)
        reps = wl.get('tile') or (1, 1, 1)
        box = np.asarray(base.box[:3], float) * 0.1
        x = np.concatenate([xb + np.asarray((i, j, k)) * box for i in range(reps[0]) for j in range(reps[1])
                            for k in range(reps[2])])
    c.set_state(wl['x'] if x is None else x)
    c.velocities_to_temperature(300.0)
    c.step(max(1, warmup))
    t0 = time.time()
    done = 0
    while done < steps and (time.time() - t0) < budget_s:
        c.step(1)
        done += 1
    dt = time.time() - t0
    return done / dt, done, dt, c.threads


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = load_workload(args.workload)
    rate, done, dt, cores = cpu_reference_run(wl, args.steps, min(args.warmup, 3), 150.0)
    line = {'metric': 'NCMC steps/s (aggregate)', 'value': rate, 'unit': 'steps/s', 'n_gpus': args.gpus, 'steps': done,
            'warmup': min(args.warmup, 3), 'ms_per_step': 1e3 * dt / done, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': wl['text'], 'replicas_per_gpu': 1},
            'cpu_baseline': {'value': rate, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d NCMC steps of one walker (time-budgeted), CPU restatement of BLUES+OpenMM '
                                       'semantics: 3 full evaluations per step, float64, not OpenMM itself' % done},
            'e2e': {'value': rate, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'ns_per_day': rate * wl['dt'] * 86.4}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--replicas', type=int, default=0, help='independent walkers per GPU (0 = the workload\'s own)')
    ap.add_argument('--workload', default='t4l', choices=sorted(WORKLOADS),
                    help='t4l = BASELINE configs[1] (default, the line the driver reads); tolparm = M1; water = M4; m5 = 250k atoms')
    ap.add_argument('--impl', default='native')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--batched', type=int, default=8, help='also report a batched run with this many walkers (0 = skip)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    from blues_b200 import mm, unit, _native
    from blues_b200.integrators import AlchemicalExternalLangevinIntegrator

    wl = load_workload(args.workload)
    system, topo, x = wl['system'], wl['topo'], wl['x']
    NSTEPS_NC, DT_PS = wl['nsteps_nc'], wl['dt']
    W = max(args.warmup, 3)
    K = args.steps
    if K + W >= NSTEPS_NC:                               # short protocols (M1): the timed region stays inside one of them
        W = min(W, max(3, NSTEPS_NC // 5))
        K = NSTEPS_NC - W - 2
    R = args.replicas or wl['replicas']
    funcs = FUNCS

    def make_context(n_rep, seed):
        integ = AlchemicalExternalLangevinIntegrator(funcs, splitting='H V R O R V H', temperature=300 * unit.kelvin,
                                                     timestep=DT_PS * unit.picoseconds, nsteps_neq=NSTEPS_NC,
                                                     nprop=1, prop_lambda=0.3)
        integ.setRandomNumberSeed(seed)
        ctx = mm.Context(system, integ, mm.Platform.getPlatformByName('CUDA'), {'DeviceIndex': local_rank},
                         n_replicas=n_rep)
        ctx.setPositions(x * unit.nanometers)
        ctx._engine.minimize(100, 10.0)                 # the surrogate force field needs a short relaxation
        ctx.setVelocitiesToTemperature(300 * unit.kelvin)
        return ctx, integ

    ctx, integ = make_context(R, 20261017 + 1000 * rank)
    eng = ctx._engine
    x_relaxed = eng.get_positions(0)                     # start of the CPU leg: same relaxed coordinates
    move = make_move(wl)
    dmove = move.device_move()
    chunk = max(1, min(100, NSTEPS_NC // 4))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    v_start = [eng.get_velocities(r) for r in range(R)]

    def restart():
        """A new protocol starts at lambda = 0 from the relaxed coordinates (as every BLUES iteration starts from an
        equilibrated MD state): restarting from the end of a cut-short protocol would switch a half-decoupled ligand
        back on inside the solvent."""
        integ.reset()
        ctx.setPositions(x_relaxed * unit.nanometers)
        for r in range(R):
            ctx.setVelocities(v_start[r] * (unit.nanometers / unit.picoseconds), replica=r)

    # ---- device-resident throughput (value) --------------------------------------------------------------------
    stream = torch.cuda.ExternalStream(eng.lib.bl_stream(eng.h))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        integ.step(W)                                    # warm-up (the sampler needs ~200 ms per sample)
        t_load = time.time()
        done = W
        while time.time() - t_load < 0.7:                # keep the GPU under the same load while clocks are sampled
            if done + 2 * chunk > NSTEPS_NC:             # a protocol is nstepsNC steps long: start the next one
                restart()
                done = 0
            integ.step(chunk)
            eng.synchronize()
            done += chunk
        # the timed region is steps W .. W+K of a fresh protocol (K + W < nstepsNC)
        restart()
        integ.step(W)
        eng.synchronize()
        barrier()
        l0 = eng.launch_count()
        e0.record(stream)
        integ.step(K)                                    # K steps, no host round-trip inside
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    if launches <= 0 or K + W >= NSTEPS_NC:
        raise SystemExit('bench: the timed region launched no kernels (K + W must stay below nstepsNC = %d)' % NSTEPS_NC)
    t = torch.tensor([ms], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * R * K / (ms_max * 1e-3)

    # ---- end to end through the public API with host buffers (e2e) ---------------------------------------------
    pos_h = [torch.from_numpy(eng.get_positions(r)).pin_memory() for r in range(R)]
    vel_h = [torch.from_numpy(eng.get_velocities(r)).pin_memory() for r in range(R)]
    # the K timed steps are the middle slice of the nstepsNC = 5000 protocol, so that the rotation move happens at
    # lambda = 0.5 (ligand fully decoupled) exactly as in a BLUES iteration (moveStep = nstepsNC / 2)
    integ.reset()
    first = max(0, NSTEPS_NC // 2 - K // 2)
    integ.setGlobalVariableByName('step', first)
    integ.setGlobalVariableByName('lambda_step', 2 * first)
    integ.setGlobalVariableByName('lambda', 2.0 * first / (2 * NSTEPS_NC))
    move_at = NSTEPS_NC // 2 - first
    barrier()
    t0 = time.perf_counter()
    for r in range(R):
        ctx.setPositions(pos_h[r].numpy() * unit.nanometers, replica=r)
        ctx.setVelocities(vel_h[r].numpy() * (unit.nanometers / unit.picoseconds), replica=r)
    if wl['move'] == 'water':
        move.beforeMove(ctx)                             # swap with a water inside the sphere (device, every walker)
    integ._scheduled_move = dict(dmove, step=move_at) if move_at < K else None
    integ.step(K)
    integ._scheduled_move = None
    if wl['move'] == 'water':
        move.afterMove(ctx)                              # out of the sphere -> protocol_work = 999999 (device)
    out_pos = [ctx.getState(getPositions=True, replica=r).getPositions(asNumpy=True) for r in range(R)]
    works = [integ.get_protocol_work(dimensionless=True, replica=r) for r in range(R)]
    acc, logp, logu = eng.accept_reject()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * R * K / float(t.item())
    nbytes = topo['n_atoms'] * 3 * 8
    h2d = 2 * nbytes * R / K
    d2h = (nbytes + 8 + 4 + 16) * R / K

    # ---- statistics gather (the only collective; outside the data path) ------------------------------------------
    from blues_b200 import parallel
    local_ids = [rank + world * r for r in range(R)]           # walker w lives on rank w % world
    gathered = parallel.gather_walker_stats(local_ids, works, logp, acc, device='cuda')
    stats = np.stack([gathered['work_kT'], gathered['accepted'].astype(float)], axis=1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel timing (roofline) on rank 0: direct launches bracketed by CUDA events ---------------------
    # restore the pre-move state (the e2e leg ended with a rotated ligand at lambda ~ 0.6; restarting the protocol at
    # lambda = 0 from there would put a fully coupled ligand on top of the protein)
    integ.reset()
    ctx.setPositions(pos_h[0].numpy() * unit.nanometers)
    ctx.setVelocities(vel_h[0].numpy() * (unit.nanometers / unit.picoseconds))
    integ.step(20)
    eng.synchronize()
    eng.set_profiling(True)
    n_prof = min(K, 200)
    integ.step(n_prof)
    eng.synchronize()
    ktimes = {}
    for name in _native.KERNEL_IDS:
        tot, n = eng.kernel_time(name)
        ktimes[name] = {'us_per_step': 1e3 * tot / n_prof, 'us_per_launch': 1e3 * tot / max(n, 1), 'launches': n}
    eng.set_profiling(False)
    pair_us = ktimes['pair']['us_per_launch']
    P_IN_PAIRS = wl['p_in']
    if P_IN_PAIRS is None:                               # pairs inside the cutoff, counted through the engine's list
        P_IN_PAIRS = int(len(eng.neighbor_pairs(0)))     # (bit-exact against the oracle's O(N^2) set in tests/)
    achieved_tflops = FLOP_PER_PAIR * P_IN_PAIRS * R / (pair_us * 1e-6) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    integ_bytes = 128 * topo['n_atoms'] * R          # 128 B/atom/launch (f64 x,v in+out; fixed-point forces in)
    integ_us = ktimes['integrate']['us_per_launch']
    # DRAM traffic per launch of the dominant kernels from the committed `ncu --set full` capture (1 walker)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
    except Exception:
        pass
    roofline = {'kernel': 'k_pair (direct-space LJ + Ewald erfc over the Verlet list, 8 lanes per atom)', 'bound': 'fp32',
                'achieved': achieved_tflops, 'peak': FP32_PEAK_NOMINAL_TFLOPS, 'unit': 'TFLOP/s',
                'frac': achieved_tflops / FP32_PEAK_NOMINAL_TFLOPS,
                'traffic': traffic.get('k_pair', {}).get('dram_bytes_per_launch'),
                'traffic_source': traffic.get('source'),
                'peak_source': 'nominal 148 SM x 128 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json has no FP32 figure; '
                               'tensor cores unused: the pair work is not a dense contraction)',
                'algorithmic_flops_per_launch': FLOP_PER_PAIR * P_IN_PAIRS * R, 'us_per_launch': pair_us}
    roofline_hbm = {'kernel': 'k_integrate (V/R/O + SHAKE/RATTLE + work bookkeeping)', 'bound': 'hbm',
                    'achieved': integ_bytes / (integ_us * 1e-6) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': integ_bytes / (integ_us * 1e-6) / 1e9 / hbm_peak,
                    'traffic': traffic.get('k_integrate', {}).get('dram_bytes_per_launch'),
                    'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650',
                    'us_per_launch': integ_us}

    # ---- batched walkers on one GPU (BASELINE configs[2] per-GPU share) -----------------------------------------
    batched = None
    if args.batched and args.batched != R and world == 1 and args.workload == 't4l':
        ctx_b, integ_b = make_context(args.batched, 777)
        nb = max(50, K // 4)
        integ_b.step(max(W // 4, 3))
        ctx_b._engine.synchronize()
        sb = torch.cuda.ExternalStream(ctx_b._engine.lib.bl_stream(ctx_b._engine.h))
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        b0.record(sb)
        integ_b.step(nb)
        b1.record(sb)
        torch.cuda.synchronize()
        bms = b0.elapsed_time(b1)
        batched = {'replicas_per_gpu': args.batched, 'steps': nb, 'value': args.batched * nb / (bms * 1e-3),
                   'unit': 'steps/s', 'ms_per_step': bms / nb}

    # ---- CPU baseline (bounded sample) ---------------------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        rate, done, dt, cores = cpu_reference_run(wl, min(40, NSTEPS_NC - 4), 2, 20.0, x=x_relaxed)
        cpu = {'value': rate, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
               'sample': '%d NCMC steps of one walker in %.1f s; CPU restatement of the reference step program (3 full '
                         'evaluations/step, float64, OpenMP) — not OpenMM itself, which is not installable here' % (done, dt)}

    line = {'metric': 'NCMC steps/s (aggregate)', 'value': value, 'unit': 'steps/s', 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': ms_max / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 pair math / f64 integration / i64 fixed-point accumulation', 'data': 'synthetic',
            'config': {'workload': wl['text'], 'replicas_per_gpu': R, 'global_walkers': world * R,
                       'l2': 'not flushed between steps: the walker state (~2 MB per walker) is the step\'s own working '
                             'set and stays L2-resident in production exactly as here',
                       'timed_region': 'K consecutive device-resident NCMC steps (CUDA-graph replay), CUDA events on the '
                                       'engine stream, max over ranks'},
            'ns_per_day': value * DT_PS * 86.4,
            'e2e': {'value': e2e_value, 'unit': 'steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'what': 'Context.setPositions/setVelocities from pinned host arrays, K steps (the slice of the protocol '
                            'around lambda = 0.5) incl. the on-device move, getState positions + protocol work '
                            '+ Metropolis test, wall clock'},
            'gpu_launches': int(launches), 'clocks': clocks.summary(), 'roofline': roofline, 'roofline_hbm': roofline_hbm,
            'kernels_us_per_step': {k: round(v['us_per_step'], 2) for k, v in ktimes.items()},
            'cpu_baseline': cpu, 'batched': batched,
            'walker_stats': {'n': int(len(stats)), 'mean_work_kT': float(stats[:, 0].mean()),
                             'accepted': int(stats[:, 1].sum())}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    # stdout carries exactly one JSON line: libraries that print to fd 1 (e.g. NCCL's version banner) go to stderr
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _out = os.fdopen(_real_stdout, 'w')
    _print = print

    def print(*a, **k):          # noqa: A001  (only the final JSON line is printed through this)
        k.setdefault('file', _out)
        _print(*a, **k)
        _out.flush()

    main()
