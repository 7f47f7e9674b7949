/* wavesim.h — C ABI of the B200-native staggered-grid FD time-stepping library.
 *
 * This is the drop-in boundary for the hot path of WAVE-Simulation (reference paths are relative to the
 * reference tree, `src/...`).  The reference has no FFI layer; its boundary is the C++ virtual interface
 *   ForwardSolver<T>::run / initForwardSolver / prepareForModelling / resetCPML   (ForwardSolver/ForwardSolver.hpp:38-52)
 * created by ForwardSolver::Factory<T>::Create(dimension,type)                   (ForwardSolver/ForwardSolverFactory.cpp:4-66)
 * and fed by Derivatives / Wavefields / Modelparameter / Acquisition objects.  The C++ host classes in
 * `wave-simulation_b200/host/` keep those names and forward to the entry points below, so a maintainer of the
 * reference binds exactly these symbols.
 *
 * Conventions: plain C; every call returns 0 on success or a negative WS_E* code, with a thread-local message
 * available from ws_last_error().  Host buffers are owned by the caller, device memory by the library.  A handle is
 * not thread-safe; handles are independent.  All calls on one handle come from one host thread (same as the
 * reference: ForwardSolver::run is non-reentrant and sequential in t).
 *
 * Linear index convention (Acquisition/Coordinates.cpp:668-694): index = x + z*NX + y*NX*NZ, x fastest,
 * y = depth (free surface at y = 0).  All arrays passed through this API use that dense, unpadded layout; the
 * padded HBM layout is private to the library.
 */
#ifndef WAVESIM_H
#define WAVESIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* equationType, as accepted by ForwardSolver::Factory::Create (ForwardSolverFactory.cpp:4-66) */
enum {
    WS_EQ_ACOUSTIC = 0,
    WS_EQ_ELASTIC = 1,
    WS_EQ_VISCOELASTIC = 2,
    WS_EQ_SH = 3,
    WS_EQ_VISCOSH = 4,
    WS_EQ_TMEM = 5,
    WS_EQ_EMEM = 6,
    WS_EQ_VISCOTMEM = 7,
    WS_EQ_VISCOEMEM = 8
};

/* source / receiver types: Acquisition/Acquisition.hpp:17-55 (SeismogramType P,VX,VY,VZ; SeismogramTypeEM EZ,EX,EY,HZ) */
enum { WS_TYPE_P = 1, WS_TYPE_VX = 2, WS_TYPE_VY = 3, WS_TYPE_VZ = 4,
       WS_TYPE_EZ = 1, WS_TYPE_EX = 2, WS_TYPE_EY = 3, WS_TYPE_HZ = 4 };

enum { WS_OK = 0, WS_EINVAL = -1, WS_ECUDA = -2, WS_ESTATE = -3, WS_ENOMEM = -4, WS_ECOMM = -5 };

#define WS_MAX_RELAX 4

/* One descriptor = the keys of the Configuration file that reach the hot path (SURVEY.md §5 "Config / flag system"). */
typedef struct ws_desc {
    int32_t dim;            /* 2 | 3                                  key: dimension                           */
    int32_t eq;             /* WS_EQ_*                                key: equationType                        */
    int32_t nx, ny, nz;     /* GLOBAL grid (nz = 1 in 2D)             keys: NX NY NZ                           */
    float dh, dt;           /*                                        keys: DH DT                              */
    int32_t nt;             /* number of time samples int(T/DT+0.5)   Simulation.cpp:304                       */
    int32_t fd_order;       /* 2,4,...,12                             key: spatialFDorder                      */
    int32_t edge_policy;    /* 0 = off-grid taps dropped (useStencilMatrix=1, Derivatives.cpp:112-121)
                               1 = order reduced towards the edge (sparse matrices, Derivatives.cpp:129-186)  */
    int32_t free_surface;   /* key FreeSurface: 0 off | 1 image method | 2 improved vacuum formulation (plain operators;
                               like 1, no absorbing frame at the top: ABS3D.cpp:197, CPML3D.cpp tests == 0)    */
    int32_t damping;        /* 0 none | 1 ABS | 2 CPML                key: DampingBoundary                     */
    int32_t boundary_width; /*                                        key: BoundaryWidth                       */
    float damping_coeff;    /*                                        key: DampingCoeff                        */
    float vmax_cpml, fc_cpml, npower; /*                              keys: VMaxCPML CenterFrequencyCPML NPower */
    int32_t n_relax;        /* L                                      key: numRelaxationMechanisms             */
    float relax_freq[WS_MAX_RELAX]; /*                                keys: relaxationFrequency[2..4]          */
    int32_t exact_arith;    /* 1 = reference operation order, no FMA contraction (bit-parity mode); 0 = FMA    */
    int32_t kernel_variant; /* 0 = auto (TMA kernels; 2-D: tile kernels), 1 = per-point kernels, 2 = cp.async marching kernels, 3 = TMA marching kernels, 4 = 2-D tile kernels */
    int32_t rank, nranks;   /* y-slab decomposition of the global grid over `nranks` processes (one GPU each)  */
    int32_t device;         /* CUDA device ordinal used by this handle                                         */
} ws_desc;

typedef struct ws_solver ws_solver; /* opaque: one per process / GPU / shot domain                             */

/* --- lifetime -------------------------------------------------------------------------------------------- */
int ws_create(const ws_desc *desc, ws_solver **out);     /* = Factory::Create + Wavefields::init + Derivatives::init */
void ws_destroy(ws_solver *s);
const char *ws_last_error(void);
const char *ws_version(void);
int ws_device_count(void);                               /* CUDA devices visible to this process (0 = none: nothing can run) */
size_t ws_estimate_memory(const ws_desc *desc);          /* bytes of HBM one rank will allocate (estimateMemory) */

/* y-range [y0, y0+nyl) of the global grid owned by this rank */
int ws_local_range(const ws_solver *s, int32_t *y0, int32_t *nyl);

/* --- model (Modelparameter getters keep their reference names) --------------------------------------------- *
 * Raw parameters : "velocityP" "velocityS" "density" "tauP" "tauS"                  (Modelparameter/ *.cpp)
 *                  "dielectricPermittivity" "electricConductivity" "magneticPermeability"
 *                  "tauDielectricPermittivity" "tauElectricConductivity"              (ModelparameterEM/ *.cpp), absolute SI
 * Derived        : "pWaveModulus" "sWaveModulus" "inverseDensity" "inverseDensityAverageX|Y|Z"
 *                  "sWaveModulusAverageXY|XZ|YZ" "tauSAverageXY|XZ|YZ" ...            (prepareForModelling products)
 * `host` holds the GLOBAL vector (n = nx*ny*nz) in linear-index order; each rank keeps its slab (+halo).       */
int ws_set_material(ws_solver *s, const char *name, const float *host, size_t n);
/* same, but `dev` is a device pointer on this handle's GPU holding this rank's slab only (nyl*nz*nx values)    */
int ws_set_material_device(ws_solver *s, const char *name, const float *dev, size_t n_local);
int ws_get_material(ws_solver *s, const char *name, float *host, size_t n); /* global gather not done: local slab, n = nyl*nz*nx */

/* Modelparameter::prepareForModelling + ForwardSolver::prepareForModelling + boundary coefficient build
 * (Elastic.cpp:28-45, ForwardSolver.cpp:34-54, CPML.cpp:39-68, ABS3D.cpp:154-218, FreeSurfaceElastic.cpp:11-47) */
int ws_prepare(ws_solver *s);

/* --- acquisition (SourceReceiverImpl, ForwardSolver/SourceReceiverImpl/SourceReceiverImpl.cpp:12-37) -------- *
 * idx1d are GLOBAL linear indices; signals is n x nt row-major (one row per source trace).                   */
int ws_set_sources(ws_solver *s, int32_t n, const int32_t *type, const int32_t *idx1d, const float *signals);
int ws_set_receivers(ws_solver *s, int32_t n, const int32_t *type, const int32_t *idx1d);
/* same with 64-bit linear indices, for global grids beyond 2^31 points (the reference's IndexType is int32, so this
 * has no reference counterpart; used by the multi-GPU weak-scaling runs) */
int ws_set_sources64(ws_solver *s, int32_t n, const int32_t *type, const int64_t *idx1d, const float *signals);
int ws_set_receivers64(ws_solver *s, int32_t n, const int32_t *type, const int64_t *idx1d);

/* --- time stepping ------------------------------------------------------------------------------------------ */
int ws_reset(ws_solver *s);                   /* Wavefields::resetWavefields + ForwardSolver::resetCPML + clear traces */
int ws_step(ws_solver *s, int32_t t);         /* = ForwardSolver::run(..., t): asynchronous enqueue of one time step   */
int ws_run(ws_solver *s, int32_t t0, int32_t t1); /* steps t0..t1-1, CUDA-graph batched                              */
int ws_sync(ws_solver *s);
/* one step through host buffers: H2D of this step's source samples (n_src floats, may be NULL = use uploaded
 * signals), step, D2H of this step's receiver samples (n_rec floats). Synchronous.                              */
int ws_step_host(ws_solver *s, int32_t t, const float *src_samples, float *rec_samples);

/* --- results -------------------------------------------------------------------------------------------------- */
/* seismogram of all receivers, n_rec x nt row-major, rows in the order given to ws_set_receivers; rows whose
 * receiver lives on another rank are left untouched (caller reduces / merges).                                  */
int ws_get_seismogram(ws_solver *s, float *host);
/* wavefield component by reference name: "VX" "VY" "VZ" "Sxx" "Syy" "Szz" "Sxy" "Sxz" "Syz" "P" "Rxx1".. /
 * "HX" "HY" "HZ" "EX" "EY" "EZ" "RX1"..  (Wavefields/Wavefields.hpp:83-141). Local slab, n = nyl*nz*nx.
 * "CURL" and "DIV": the snapType 3 energy measures computed from the particle velocities (elastic / viscoelastic,
 * Wavefields3Delastic.cpp:197-245 getCurl / getDiv) or the magnetic field (2-D TMEz, 3-D EM: WavefieldsEM/).    */
int ws_get_wavefield(ws_solver *s, const char *comp, float *host, size_t n);
int ws_set_wavefield(ws_solver *s, const char *comp, const float *host, size_t n);
int ws_is_finite(ws_solver *s, int32_t *flag); /* Wavefields::isFinite + SeismogramHandler::isFinite, Simulation.cpp:519 */

/* --- wavefield objects and their operators (hooks of WAVE-Inversion) ------------------------------------- *
 * The reference's time loop and the inversion code that links libSimulation work on whole Wavefields objects: a copy taken
 * before a step (`*wavefieldsTemp = *wavefields`, Simulation.cpp:450), `-=`, `+=`, `*= scalar`, `*= vector`
 * (Wavefields/Wavefields.hpp:62-80; every concrete class applies the operator to all of its components incl. the memory
 * variables, e.g. Wavefields3Dviscoelastic.cpp:305-330).  A ws_wavefields is such a second set of components, resident in HBM in
 * the layout of the handle it was created from; NULL stands for the solver's own (live) wavefields.  Element-wise fp32
 * operations, one rounding each, asynchronous on the handle's stream.                                                          */
typedef struct ws_wavefields ws_wavefields;
int ws_wavefields_create(ws_solver *s, ws_wavefields **out);  /* all components zero (Wavefields::init)                        */
void ws_wavefields_destroy(ws_wavefields *w);
int ws_wavefields_assign(ws_solver *s, ws_wavefields *dst, const ws_wavefields *src);        /* operator=                       */
int ws_wavefields_plus_assign(ws_solver *s, ws_wavefields *dst, const ws_wavefields *src);   /* operator+=                      */
int ws_wavefields_minus_assign(ws_solver *s, ws_wavefields *dst, const ws_wavefields *src);  /* operator-=                      */
int ws_wavefields_times_assign(ws_solver *s, ws_wavefields *dst, float rhs);                 /* operator*=(ValueType)           */
/* operator*=(DenseVector): `host` holds this rank's slab of the vector (n = nyl*nz*nx values, like ws_set_wavefield)           */
int ws_wavefields_times_assign_vector(ws_solver *s, ws_wavefields *dst, const float *host, size_t n);
int ws_wavefields_get(ws_solver *s, const ws_wavefields *w, const char *comp, float *host, size_t n); /* like ws_get_wavefield */
/* `*wavefields *= compensation` after every time step (Simulation.cpp:455-456; Modelparameter::getCompensation, EM only): the
 * vector stays in HBM and the multiplication rides in the captured step graph.  host = NULL switches it off.                    */
int ws_set_step_scaling(ws_solver *s, const float *host, size_t n);

/* --- multi-GPU (replaces src/Partitioning; y-slab decomposition) --------------------------------- *
 * One process per GPU.  Rank 0 creates an id with ws_comm_unique_id, the launcher broadcasts the 128 bytes, every
 * rank calls ws_comm_init.  Halo planes are exchanged with ncclSend/ncclRecv on a communication stream overlapped
 * with the interior kernels.                                                                                      */
int ws_comm_unique_id(void *id128);
int ws_comm_init(ws_solver *s, const void *id128);
/* Where every neighbour's memory can be mapped (GPUs of one NVLink / NVSwitch node: peer access inside a process, CUDA IPC between
 * processes) ws_comm_init switches the halo planes of the time loop from ncclSend / ncclRecv to the library's own kernels: the sender
 * stores its edge planes straight into the neighbour's ghost planes and raises a counter there, the neighbour's edge-slab kernels wait
 * for the counter.  NCCL stays for the collectives (isFinite, set-up exchanges).  env WS_P2P=0 keeps NCCL for everything.
 * 0 = single rank, 1 = NCCL send / recv, 2 = external transport, 3 = kernels over peer memory                                     */
int ws_halo_transport(const ws_solver *s);
/* Bring-your-own transport instead of NCCL (e.g. CUDA-aware MPI, the dmemo::Communicator role; gloo in the CPU tests):
 * `fn` must send `count` floats from `send` to rank `peer` and receive `count` floats from it into `recv` (device
 * pointers on this handle's GPU), returning 0 on success.  Synchronous: the library drains its streams before each
 * call, so there is no compute/communication overlap on this path.                                                */
typedef int (*ws_sendrecv_fn)(void *user, const float *send, float *recv, size_t count, int peer);
int ws_comm_init_external(ws_solver *s, ws_sendrecv_fn fn, void *user);

/* --- operator-given mode: irregular grids ------------------------------------------------------------- *
 * On a variable grid (layers of spacing 3^n DH, Acquisition/Coordinates.cpp:115-247) or with a variable FD order the reference's
 * derivative matrices are no stencils: it assembles them point by point (ForwardSolver/Derivatives/Derivatives.cpp:129-1243) together
 * with interpolation matrices for the interface planes (:1252-1566), and its boundary profiles are per-point sparse vectors
 * (BoundaryCondition/CPML2DAcoustic.cpp:99-200).  The host layer (wave-simulation_b200/host/) assembles the same rows and hands
 * them over here; the library then runs the reference's statement sequence (ForwardSolver{2D,3D}acoustic.cpp run()) with one
 * fused gather kernel per half-step.  Acoustic solvers; one GPU per shot.  All vectors hold n_points values in the model's own
 * order (Coordinates::index2coordinate), source / receiver indices are model-vector indices.
 * ws_create_sparse replaces ws_create; the model parameters the kernels read ("pWaveModulus", "inverseDensityAverageX|Y|Z") are
 * given with ws_set_material; ws_prepare only checks that everything is there.                                                   */
int ws_create_sparse(const ws_desc *desc, int64_t n_points, ws_solver **out);
/* operator rows in ELL form: cols / vals are n_points x max_taps row-major, columns ascending, unused entries cols = -1.
 * name: "Dxf" "Dxb" "Dyf" "Dyb" "Dzf" "Dzb"; values already carry DT (FDTD2D.cpp:213-216).  With FreeSurface = 1 the acoustic
 * solvers use DyfFreeSurface in place of Dyf (ForwardSolver2Dacoustic.cpp:141-146): pass that matrix as "Dyf".                    */
int ws_set_operator(ws_solver *s, const char *name, int32_t max_taps, const int32_t *cols, const float *vals);
/* interpolation matrices "InterpolationFull" "InterpolationStaggeredX" "InterpolationStaggeredZ": only the n_rows rows that are not
 * identity rows (the points of the interface planes), rows[] = their model-vector indices.                                        */
int ws_set_interpolation(ws_solver *s, const char *name, int64_t n_rows, const int32_t *rows, int32_t max_taps, const int32_t *cols, const float *vals);
/* CPML coefficients of one axis (0 x, 1 y, 2 z) as the reference's sparse vectors a, b, a_half, b_half on the points idx[]        */
int ws_set_cpml_profile(ws_solver *s, int32_t axis, int64_t n, const int32_t *idx, const float *a, const float *b, const float *a_half, const float *b_half);
/* ABS frame (DampingBoundary = 1) as the reference's sparse vector `damping` (ABS2D.cpp:110-178, ABS3D.cpp:154-218): the wavefields
 * are multiplied by damping[k] on the points idx[k] after the pressure update                                                   */
int ws_set_abs_profile(ws_solver *s, int64_t n, const int32_t *idx, const float *damping);
/* points of the free surface (FreeSurface::setSurfaceZero, FreeSurface.cpp:13-20)                                                 */
int ws_set_surface(ws_solver *s, int64_t n, const int32_t *idx);

/* --- instrumentation --------------------------------------------------------------------------------------------- */
/* number of kernel launches issued by this handle since creation */
uint64_t ws_launch_count(const ws_solver *s);
/* average device time (ms) of the dominant kernels of the last ws_run (CUDA events on the compute stream):
 * which = 0 first half-step kernel (velocity / H), 1 second half-step kernel (stress / E), 2 whole step         */
int ws_last_timing(ws_solver *s, int which, float *ms);
void *ws_stream(ws_solver *s); /* cudaStream_t the kernels are launched on */
/* enable = 1: ws_run launches directly and brackets both half-step kernels of every step with CUDA events */
int ws_set_timing(ws_solver *s, int enable);
/* 1 if the tiled TMA kernels (not the general per-point kernels) serve this configuration */
int ws_uses_fast_kernels(const ws_solver *s);
/* which kernel family serves this configuration: 0 per-point (general), 1 marching (register queue + cp.async-staged
 * planes), 2 warp-specialised TMA kernels of the 3-D elastic case, 3 warp-specialised TMA marching kernels (all solvers),
 * 4 2-D tile kernels (one thread block per 128 x 8 tile, operands by TMA onto one mbarrier; all 2-D solvers) */
int ws_kernel_path(const ws_solver *s);

#ifdef __cplusplus
}
#endif
#endif /* WAVESIM_H */
