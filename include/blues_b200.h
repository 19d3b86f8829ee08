/*
 * blues_b200 — C ABI of the B200-native NCMC engine (libblues_b200.so).
 *
 * The reference (MobleyLab/blues) has no FFI of its own: its hot path crosses from Python into OpenMM
 * through SWIG on every Simulation.step(1) / Context.getState / setPositions / get|setGlobalVariableByName
 * call.  Each entry point below replaces one of those crossings; the `replaces:` note cites the reference
 * call site (file:line under the reference checkout).  INTEGRATION.md shows the ctypes binding a BLUES
 * maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns 0 on success or a negative bl_status; bl_last_error() returns the message.
 *   - one host thread per handle; a handle is bound to one CUDA device and owns one stream.
 *   - host buffers are caller-owned, row-major, float64; units nm, ps, dalton, kJ/mol, e, K, radian.
 *   - `replica` selects one of the n_replicas independent walkers held by the handle; -1 = all
 *     (set: broadcast the same host data; get: not allowed unless stated).
 */
#ifndef BLUES_B200_H
#define BLUES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bl_handle bl_handle;

typedef enum bl_status {
    BL_OK = 0,
    BL_ERR_INVALID = -1,      /* bad argument / unsupported topology feature                       */
    BL_ERR_CUDA = -2,         /* CUDA runtime / cuFFT failure                                      */
    BL_ERR_NAN = -3,          /* a walker produced non-finite coordinates (OpenMM raises likewise) */
    BL_ERR_CAPACITY = -4,     /* neighbour-tile list overflow (retry after bl_reserve_tiles)       */
    BL_ERR_NO_DEVICE = -5     /* no CUDA device: the engine has no CPU fallback                    */
} bl_status;

/* Flat system description = what parmed.Structure.createSystem + openmmtools' AbsoluteAlchemicalFactory hand to
 * OpenMM in the reference (blues/simulation.py:139-219, 221-317).  Arrays are copied by bl_create. */
typedef struct bl_topology {
    int32_t n_atoms;
    const double* mass;            /* [n_atoms] dalton; 0 = frozen (blues/utils.py:202-221)        */
    const double* charge;          /* NonbondedForce per-particle q (alchemical atoms: 0)          */
    const double* sigma;           /* nm                                                            */
    const double* epsilon;         /* kJ/mol (alchemical atoms: 0)                                  */
    int32_t n_bonds;     const int32_t* bonds;     const double* bond_k;    const double* bond_r0;
    int32_t n_angles;    const int32_t* angles;    const double* angle_k;   const double* angle_t0;
    int32_t n_torsions;  const int32_t* torsions;  const double* torsion_k; const int32_t* torsion_n;
    const double* torsion_phase;
    int32_t n_excl;      const int32_t* excl_pairs; /* [n_excl][2] i<j: exclusions AND exceptions   */
    const double* excl_qq; const double* excl_sigma; const double* excl_eps; /* 0,·,0 = plain exclusion */
    int32_t n_constraints; const int32_t* constraints; const double* constraint_d;
    double  box[3];                /* orthorhombic box edge lengths                                 */
    int32_t nb_method;             /* 0 NoCutoff, 2 CutoffPeriodic (reaction field), 4 PME           */
    double  cutoff;
    double  ewald_alpha;
    int32_t pme_grid[3];
    double  dispersion_coeff;      /* E_disp = coeff / V                                             */
    int32_t remove_cm;             /* CMMotionRemover every step                                     */
    int32_t n_restraints; const int32_t* restraint_atoms; const double* restraint_k; const double* restraint_x0;
    /* alchemical region (openmmtools softcore) */
    int32_t n_alch;      const int32_t* alch_atoms; const double* alch_charge; const double* alch_sigma;
    const double* alch_eps;
    int32_t n_alch_exc;  const int32_t* alch_exc_pairs; const double* alch_exc_qq; const double* alch_exc_sigma;
    const double* alch_exc_eps;
    double  softcore_alpha, softcore_a, softcore_b, softcore_c;
    int32_t annihilate_sterics, annihilate_electrostatics;
    /* Generic Custom*Force terms (OpenMM CustomNonbondedForce over interaction groups or all pairs, CustomBondForce,
       two-group CustomCentroidBondForce — blues/tests/data/ethylene_system.xml:52-114 is the reference's use):
       E = f(r; per-term parameters, lambda_sterics, lambda_electrostatics), r = distance between the weighted
       centroids of two atom groups.  f arrives as a stack program compiled by the host from the Lepton expression
       (blues_b200/lepton.py::compile_program); the device evaluates value and d/dr with forward-mode dual numbers.
       The two lambdas follow the integrator's tables exactly like the softcore terms. */
    int32_t n_custom_terms;
    const int32_t* custom_term;          /* [n_custom_terms][4]: group A, group B, program, flags (bit 0: minimum image) */
    const double*  custom_cutoff;        /* [n_custom_terms] nm; <= 0: no cutoff                                        */
    int32_t custom_n_params;             /* parameters per term                                                         */
    const double*  custom_params;        /* [n_custom_terms][custom_n_params]                                           */
    int32_t n_custom_groups;
    const int32_t* custom_group_start;   /* [n_custom_groups + 1] offsets into the two arrays below                     */
    const int32_t* custom_group_atoms;
    const double*  custom_group_weights; /* normalised (sum 1 per group)                                                */
    int32_t n_custom_progs;
    const int32_t* custom_prog_start;    /* [n_custom_progs + 1] offsets into the code arrays                           */
    const int32_t* custom_code_op;       /* BL_OP_* */
    const double*  custom_code_arg;      /* immediate: constant, parameter / global index, integer exponent             */
} bl_topology;

/* stack-program opcodes of the custom-force evaluator */
#define BL_OP_CONST   0
#define BL_OP_R       1   /* the distance */
#define BL_OP_PARAM   2   /* per-term parameter arg */
#define BL_OP_GLOBAL  3   /* 0 lambda_sterics, 1 lambda_electrostatics */
#define BL_OP_ADD     4
#define BL_OP_SUB     5
#define BL_OP_MUL     6
#define BL_OP_DIV     7
#define BL_OP_NEG     8
#define BL_OP_POWI    9   /* x^arg, arg integer */
#define BL_OP_POW    10   /* a^b, both from the stack */
#define BL_OP_SQRT   11
#define BL_OP_EXP    12
#define BL_OP_LOG    13
#define BL_OP_SIN    14
#define BL_OP_COS    15
#define BL_OP_TAN    16
#define BL_OP_ABS    17
#define BL_OP_MIN    18
#define BL_OP_MAX    19
#define BL_OP_STEP   20
#define BL_OP_DELTA  21
#define BL_OP_SELECT 22   /* select(c, a, b): a if c != 0 else b */
#define BL_OP_ERF    23
#define BL_OP_ERFC   24
#define BL_OP_TANH   25
#define BL_OP_SINH   26
#define BL_OP_COSH   27
#define BL_OP_ATAN   28
#define BL_CUSTOM_STACK 24

/* Integrator kinds */
#define BL_INTEGRATOR_NCMC      1   /* AlchemicalExternalLangevinIntegrator (blues/integrators.py:8-249)      */
#define BL_INTEGRATOR_LANGEVIN  2   /* openmm.LangevinIntegrator for the MD leg (blues/simulation.py:628-648) */

typedef struct bl_integrator_params {
    int32_t kind;
    double  temperature;        /* K                                                                  */
    double  friction;           /* 1/ps (collision_rate)                                              */
    double  timestep;           /* ps                                                                 */
    double  constraint_tol;     /* relative, blues/integrators.py:104 default 1e-8                    */
    const char* splitting;      /* e.g. "H V R O R V H" (NCMC only)                                   */
    int32_t nsteps_neq;         /* NCMC only                                                          */
    int32_t nprop;              /* blues/integrators.py:108                                           */
    double  prop_lambda_min, prop_lambda_max;   /* blues/integrators.py:147-157                       */
    int32_t n_lambda;           /* length of the two tables = nsteps_neq * n_H + 1                    */
    const double* lambda_sterics;          /* alchemical_functions evaluated at lambda_step/n_lambda_steps */
    const double* lambda_electrostatics;
} bl_integrator_params;

/* On-device moves (blues/moves.py) */
#define BL_MOVE_NONE             0
#define BL_MOVE_ROTATE           1   /* RandomLigandRotationMove.move, blues/moves.py:278-310              */
#define BL_MOVE_WATER_SWAP       2   /* WaterTranslationMove.beforeMove, blues/moves.py:951-1007: a uniformly
                                        chosen water whose first atom lies within `radius` (periodic, float32)
                                        of the centre of mass of `center_atoms` trades positions and velocities
                                        with the alchemical water `atoms`; none found -> the walker's move is off */
#define BL_MOVE_WATER_TRANSLATE  3   /* WaterTranslationMove.move, blues/moves.py:1009-1053: alchemical water
                                        translated so that its first atom sits on a uniform point of the sphere */
#define BL_MOVE_WATER_CHECK      4   /* WaterTranslationMove.afterMove, blues/moves.py:1055-1083: first atom
                                        outside the sphere -> protocol_work = 999999 (forced rejection)         */
typedef struct bl_move {
    int32_t kind;
    int32_t step;               /* moveStep: applied before the integrator step with this index       */
    int32_t n_atoms;            /* ROTATE: ligand atoms; WATER_*: atoms of the alchemical water        */
    const int32_t* atoms;
    const double* masses;       /* ROTATE: element masses used for the centre of mass (float32 arithmetic); WATER_*: unused */
    /* WATER_* only */
    int32_t n_waters;           /* candidate water residues (WATER_SWAP)                               */
    const int32_t* water_atoms; /* [n_waters][n_atoms] atom indices, first atom of a row = its oxygen   */
    int32_t n_center;           /* atoms of the protein selection defining the sphere centre           */
    const int32_t* center_atoms;
    const double* center_masses;/* element masses (float32 arithmetic)                                 */
    double radius;              /* nm                                                                  */
} bl_move;

/* ---- lifecycle ------------------------------------------------------------------------------------- */
/* replaces: openmm.app.Simulation(topology, system, integrator, platform) — blues/simulation.py:730-737 */
int bl_create(const bl_topology* topo, int device, int n_replicas, uint64_t seed, bl_handle** out);
int bl_destroy(bl_handle* h);
const char* bl_last_error(const bl_handle* h);     /* h may be NULL: error of the last failed bl_create */
int bl_num_replicas(const bl_handle* h);
int bl_num_atoms(const bl_handle* h);
/* replaces: integrator construction — blues/simulation.py:628-705, blues/integrators.py:98-145 */
int bl_set_integrator(bl_handle* h, const bl_integrator_params* p);
int bl_set_seed(bl_handle* h, uint64_t seed);                    /* integrator.setRandomNumberSeed */

/* ---- state in / out ---------------------------------------------------------------------------------- */
/* replaces: context.setPositions / setVelocities / setPeriodicBoxVectors — blues/simulation.py:956-962 */
int bl_set_positions(bl_handle* h, int replica, const double* xyz);
int bl_set_velocities(bl_handle* h, int replica, const double* vxyz);
int bl_set_box(bl_handle* h, const double box[3]);
/* replaces: context.getState(getPositions, getVelocities, getForces, getEnergy) — blues/simulation.py:905 */
int bl_get_positions(bl_handle* h, int replica, double* xyz);
int bl_get_velocities(bl_handle* h, int replica, double* vxyz);
int bl_get_forces(bl_handle* h, int replica, double* fxyz);
int bl_get_box(bl_handle* h, double box[3]);
int bl_get_energy(bl_handle* h, double* epot /*[R]*/, double* ekin /*[R]*/);
#define BL_NUM_ENERGY_TERMS 12
/* order: bond, angle, torsion, restraint, lj+coulomb direct, exceptions+ewald exclusion, pme reciprocal,
 *        ewald self+plasma, dispersion, alchemical sterics, alchemical electrostatics, alchemical exceptions */
int bl_get_energy_terms(bl_handle* h, int replica, double terms[BL_NUM_ENERGY_TERMS]);
/* device-to-device copy of box/positions/velocities between two handles on the same device
 * replaces: _syncStatesMDtoNCMC — blues/simulation.py:1028-1037 (flags: 1 positions, 2 velocities, 4 box) */
int bl_copy_state(bl_handle* dst, const bl_handle* src, int flags);
/* the same for the walkers r with mask[r] != 0 only (mask NULL: all) — the accepted walkers of a many-walker
 * _acceptRejectMove take the NCMC end positions, blues/simulation.py:1142-1146 */
int bl_copy_state_masked(bl_handle* dst, const bl_handle* src, int flags, const int32_t* mask /*[R]*/);
/* replaces: context.setVelocitiesToTemperature(T) — blues/simulation.py:743,1187 */
int bl_velocities_to_temperature(bl_handle* h, double temperature);

/* ---- integrator globals ---------------------------------------------------------------------------- */
/* replaces: integrator.get/setGlobalVariableByName — blues/simulation.py:872, blues/integrators.py:233-249,
 * blues/moves.py:1082.  Names: lambda, lambda_step, step, protocol_work, shadow_work, heat, first_step,
 * perturbed_pe, unperturbed_pe, prop, nprop, prop_lambda_min, prop_lambda_max, Eold, Enew, debug,
 * lambda_sterics, lambda_electrostatics, n_lambda_steps, nsteps, kT. */
int bl_get_global(bl_handle* h, int replica, const char* name, double* value);
int bl_set_global(bl_handle* h, int replica, const char* name, double value);
/* replaces: AlchemicalExternalLangevinIntegrator.reset() — blues/integrators.py:240-249 */
int bl_reset_ncmc(bl_handle* h);

/* ---- the hot path ---------------------------------------------------------------------------------- */
/* Advance every walker n_steps NCMC steps on the device without host round-trips; optional on-device move.
 * replaces: the `for step in range(nstepsNC)` loop of BLUESSimulation._stepNCMC — blues/simulation.py:1066-1094 */
int bl_ncmc_run(bl_handle* h, int n_steps, const bl_move* move /* nullable */);
/* replaces: _stepMD loop — blues/simulation.py:1203-1205 */
int bl_md_run(bl_handle* h, int n_steps);
/* apply a move now (between integrator steps); the next NCMC step accounts its work as external work
 * replaces: MoveEngine.runEngine(context) — blues/moves.py:385-410 */
int bl_apply_move(bl_handle* h, const bl_move* move);
/* Metropolis test per walker on the device: accept iff logp + correction > log(u)
 * replaces: _acceptRejectMove — blues/simulation.py:1121-1146.  correction may be NULL (treated as 0). */
int bl_accept_reject(bl_handle* h, const double* correction, int32_t* accepted, double* logp, double* log_u);
/* replaces: simulation.minimizeEnergy(maxIterations) — blues/tests/test_simulation.py:139-141 */
int bl_minimize(bl_handle* h, int max_iterations, double tolerance);

/* ---- introspection (tests, bench) ---------------------------------------------------------------- */
/* sorted codes i*n_atoms+j (i<j) of the non-excluded pairs within the cutoff at the current positions */
int bl_neighbor_pairs(bl_handle* h, int replica, int64_t* codes, size_t capacity, size_t* n_pairs);
int bl_neighbor_stats(bl_handle* h, int replica, int64_t* n_tiles, int64_t* n_rebuilds);
uint64_t bl_launch_count(const bl_handle* h);      /* kernels launched by this handle so far            */
/* per-kernel CUDA-event timing of direct (non-graph) launches; kernel ids below */
#define BL_K_PAIR 0
#define BL_K_INTEGRATE 1
#define BL_K_PME_SPREAD 2
#define BL_K_PME_GATHER 3
#define BL_K_PME_CONVOLVE 4
#define BL_K_BONDED 5
#define BL_K_ALCH 6
#define BL_K_NEIGHBOR 7
#define BL_K_FFT 8
#define BL_NUM_KERNEL_IDS 9
int bl_set_profiling(bl_handle* h, int on);
int bl_get_kernel_time(bl_handle* h, int kernel_id, double* total_ms, int64_t* launches);
int bl_use_graphs(bl_handle* h, int on);           /* CUDA-graph replay of the step program (default on) */
void* bl_stream(bl_handle* h);                     /* cudaStream_t the handle launches on                */
int bl_synchronize(bl_handle* h);
const char* bl_version(void);
/* FP32 FMA throughput of the device measured by a register-only microbenchmark (TFLOP/s, best of 10 launches):
 * the roofline denominator bench.py reports the pair kernel against. */
int bl_measure_fp32_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* BLUES_B200_H */
