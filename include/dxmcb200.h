/* dxmcb200.h — C ABI of the B200 photon-transport runtime (libdxmcb200.so).
 *
 * This is the seam the reference crosses at Transport<T>::parallellRun
 * (reference include/dxmc/transport.hpp:765-778): everything the reference's worker threads read
 * (world arrays, attenuation look-up tables, beam tables, exposures) goes in as plain data,
 * the three per-voxel result arrays come out. POD only, no C++ or torch types, no exceptions;
 * every function returns 0 on success or a negative dxmcb200_status.
 *
 * Callers: the drop-in C++ classes in dxmclib_b200/include/dxmc/ (Transport<T>::operator()),
 * and any FFI (ctypes in tests/ and bench.py). There is NO CPU fallback: without a CUDA device
 * dxmcb200_create fails with DXMCB200_ERR_NO_DEVICE.
 *
 * Threading: one ctx per device per host thread. dxmcb200_run is synchronous for the caller and
 * asynchronous on the ctx's own CUDA streams inside (two wave pipelines). dxmcb200_set_world and
 * dxmcb200_get_result move the caller's pageable arrays through pinned staging buffers on a few helper
 * threads of their own; large device blocks are pooled per process (dxmcb200_trim_pool).
 */
#ifndef DXMCB200_H
#define DXMCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dxmcb200_ctx dxmcb200_ctx;

enum dxmcb200_status {
    DXMCB200_OK = 0,
    DXMCB200_ERR_ARG = -1,
    DXMCB200_ERR_NO_DEVICE = -2,
    DXMCB200_ERR_CUDA = -3,
    DXMCB200_ERR_STATE = -4,
    DXMCB200_ERR_CANCELLED = -5,
    DXMCB200_ERR_NCCL = -6
};

/* ---- inputs ----------------------------------------------------------------------------- */

/* World<T> as the hot loop sees it (reference world.hpp:37-68, transport.hpp:485-521, 643-645).
 * Voxel order is x fastest: idx = z*nx*ny + y*nx + x (transport.hpp:514). Arrays are HOST pointers
 * and are copied (packed on the device: one 8-byte record per voxel, or one byte / half a byte per voxel
 * plus a 256-entry record table when the grid holds at most 256 / 16 distinct records). measurement may be NULL. */
typedef struct dxmcb200_world {
    uint64_t dim[3];
    float spacing[3];
    float extent_safe[6]; /* World::matrixExtentSafe(): x0 x1 y0 y1 z0 z1 */
    const float* density;
    const uint8_t* material;
    const uint8_t* measurement;
} dxmcb200_world;

enum { DXMCB200_RITA_N = 56, DXMCB200_SPLINE_N = 16, DXMCB200_SHELLS = 12, DXMCB200_SHELL_FLOATS = 11,
    DXMCB200_SPLINE_FLOATS = 63 };
/* tracking mode 1: an air walk crosses at most this many all-air cubes; a photon that is still in air then rejoins the Woodcock
 * steps where it stands (and is picked up by the next walk). Bounds the trip count of the walk loop, whose slowest lane sets a
 * warp's pace; shared with the CPU restatement so that both walk draw for draw alike. */
#ifndef DXMCB200_WALK_MAX_CUBES_VALUE
#define DXMCB200_WALK_MAX_CUBES_VALUE 3
#endif
enum { DXMCB200_WALK_MAX_CUBES = DXMCB200_WALK_MAX_CUBES_VALUE };

/* AttenuationLut<T> flattened (reference attenuationlut.hpp:267-274, attenuationinterpolator.hpp:37-45).
 * All arrays are HOST pointers, copied. */
typedef struct dxmcb200_luts {
    uint32_t n_materials;
    uint32_t n_segments;        /* m_resolution after generate(): number of log-log segments */
    uint32_t linear_index;      /* m_linearIndex */
    float linear_step;          /* m_linearStep */
    float linear_energy;        /* m_linearEnergy (log10 keV) */
    const float* knots;         /* m_x: [n_segments] upper segment edges, log10 keV */
    const float* coefficients;  /* m_coefficients: [n_materials][n_segments][photo,compton,rayleigh][b,a] */
    const float* max_coefficients; /* m_maxCoefficients: [n_segments][b,a] of the Woodcock majorant inverse */
    const float* rita;          /* form-factor sampler RITA<T,56>: [n_materials][x|e|a|b][56] (dxmcrandom.hpp:334-339) */
    const float* spline;        /* scatter function CubicSplineInterpolator<T,16>: [n_materials][60 coeffs, start, step, stop] */
    const float* shells;        /* ElectronShellConfiguration<T>[12] per material: [n_materials][12]
                                   {binding, nElectrons, J0, photoIonProb, yield, lineProb[3], lineEnergy[3]} */
} dxmcb200_luts;

/* SpecterDistribution<T> (alias table, reference dxmcrandom.hpp:173-330) */
typedef struct dxmcb200_spectrum {
    uint32_t n;
    const float* probs;    /* m_probs */
    const uint32_t* alias; /* m_alias */
    const float* energies; /* m_energies */
} dxmcb200_spectrum;

/* HeelFilter<T> (reference beamfilters.hpp:436-538) */
typedef struct dxmcb200_heel {
    float energy_start, energy_step;
    uint32_t energy_size;
    float angle_start, angle_step;
    uint32_t angle_size;
    const float* weights; /* [energy_size][angle_size] */
} dxmcb200_heel;

/* BowTieFilter<T> (reference beamfilters.hpp:78-199): sorted |angle|, normalised weight */
typedef struct dxmcb200_bowtie {
    uint32_t n;
    const float* angles;
    const float* weights;
} dxmcb200_bowtie;

/* Exposure<T> after alignToDirectionCosines (reference exposure.hpp:36-304) */
typedef struct dxmcb200_exposure {
    float position[3];
    float cosines[6];
    float beam_direction[3];
    float collimation[4]; /* x0 x1 y0 y1 */
    float weight;         /* m_beamIntensityWeight */
    float mono_energy;    /* used when spectrum < 0 */
    int32_t spectrum;     /* index into the spectra table or -1 */
    int32_t heel;         /* index into the heel table or -1 */
    int32_t bowtie;       /* index into the bowtie table or -1 */
    uint32_t reserved;
    uint64_t histories;
} dxmcb200_exposure;

/* ---- tube spectrum ------------------------------------------------------ */
/* Thick-target tungsten bremsstrahlung of Poludniowski & Evans as the reference evaluates it per energy bin and take-off angle
 * (betheHeitlerCrossSection.hpp:368-407 from tube.hpp:191-208 and beamfilters.hpp:465-505: 141 depths x 200 electron energies per
 * bin, 0.4 s of host time for a CT source with heel model): out[a * n_bins + b] = betheHeitlerSpectra(tube_voltage, energies[b],
 * angles[a]) on the calling thread's current CUDA device. tungsten_attenuation[b] = Material::getTotalAttenuation(74, energies[b])
 * from the host's element data. Same single-precision operations in the same order as the host code, CUDA's libm instead of the
 * host's (2e-5 relative on a normalised spectrum). Tube::getSpecter uses it when DXMCB200_DEVICE_SPECTRUM=1; the default is the
 * host path, whose tables are bit-identical to the reference's. */
int dxmcb200_tube_bremsstrahlung(float tube_voltage, uint32_t n_bins, const float* energies, const float* tungsten_attenuation, uint32_t n_angles,
    const float* angles, float* out);

/* ---- physics data ----------------------------------------------------------------------- */
/* Name of the element data source the host-side table builders (Material, AttenuationLut, Tube) of this library were
 * compiled against: "xraylib" (what the reference uses, src/material.cpp:23-375) or the in-repo "xrl_lite (approximate ...)".
 * *approximate (may be NULL) is 1 for the latter: absolute doses then carry the few-% error of that data (TG-195 totals
 * within about 2-5 %); product-vs-reference parity is unaffected because both are built on the same source here. The
 * library also prints this line once to stderr when approximate data is first used (DXMCB200_QUIET=1 silences it). */
const char* dxmcb200_physics_backend(int* approximate);

/* A whole source as one parameter block: what exposure i looks like is a pure function of (block, i), evaluated for all i
 * on the device by dxmcb200_generate_exposures (replaces the reference's per-exposure virtual call Source::getExposure(i),
 * source.hpp:102, for every source type it has; the formulas and their reference lines are in
 * dxmclib_b200/include/dxmc/sourcemodel.hpp, whose model::SourceParams<float> has exactly this layout). */
enum dxmcb200_source_motion {
    DXMCB200_SOURCE_FIXED = 0,           /* every exposure identical: pencil, isotropic, DX */
    DXMCB200_SOURCE_ORBIT = 1,           /* focal spot and frame turn about an axis: isotropic CT, cone-beam CT */
    DXMCB200_SOURCE_GANTRY_AXIAL = 2,    /* CT gantry, step-and-shoot */
    DXMCB200_SOURCE_GANTRY_SPIRAL = 3,   /* CT gantry, continuous table feed */
    DXMCB200_SOURCE_GANTRY_TOPOGRAM = 4  /* CT gantry parked, table moving */
};
typedef struct dxmcb200_source_params {
    uint32_t motion;
    uint32_t tubes;            /* gantry: 2 = dual source, even exposure indices tube A, odd tube B */
    uint64_t exposures;
    uint64_t histories;        /* per exposure */
    float position[3];         /* Source::position() */
    float cosines[6];          /* Source::directionCosines() */
    float collimation[4];      /* FIXED / ORBIT: x0 x1 y0 y1 */
    float mono_energy;         /* used when spectrum[0] < 0 */
    float focal_offset;        /* FIXED / ORBIT: focal spot = position - beam * focal_offset */
    int32_t spectrum[2], heel[2], bowtie[2]; /* beam-table indices per tube, -1 = none */
    uint32_t orbit_full_turn;  /* 1: angle_i = 2 pi i / exposures about z through the origin; 0: i * orbit_step about the y cosine through position */
    float orbit_step;
    float sdd[2], fov[2], start_angle[2], tube_weight[2]; /* gantry, per tube */
    float beam_width;          /* collimation along z at the isocentre [mm] */
    float angle_step, pitch, table_step, tilt, scan_length;
    uint32_t xcare;            /* organ-based tube current modulation on */
    float xcare_angle, xcare_span, xcare_ramp, xcare_low;
    uint32_t aec_size;         /* entries of the tube-current profile along z handed to dxmcb200_generate_exposures (0: none) */
    float aec_min, aec_max, aec_step;
    uint32_t align;            /* 1: express the exposures in the basis of world_cosines (Exposure::alignToDirectionCosines) */
    float world_cosines[6];
} dxmcb200_source_params;

/* ---- life cycle ------------------------------------------------------------------------- */
int dxmcb200_device_count(int* count);
int dxmcb200_create(int device, dxmcb200_ctx** out);
void dxmcb200_destroy(dxmcb200_ctx*);
const char* dxmcb200_last_error(dxmcb200_ctx*);

int dxmcb200_set_world(dxmcb200_ctx*, const dxmcb200_world*);
/* Per-material maximum density of the uploaded grid, out[m] for m < n_materials (<= 256): the one whole-grid
 * quantity the Woodcock majorant needs (replaces the per-material transform_reduce over all voxels of
 * attenuationinterpolator.hpp:48-59 with one device pass, or a look at the 256-entry table of a palette grid). */
int dxmcb200_material_max_density(dxmcb200_ctx*, uint32_t n_materials, float* out);
/* Large device blocks (accumulators, wave buffers, staging arrays) can be parked in a per-process pool when a context is
 * destroyed and reused by the next one (saves the cudaMalloc / cudaFree of tens of GB per Transport call). OFF by default:
 * nothing stays allocated after a context is destroyed. dxmcb200_set_pool_limit(bytes) (or DXMCB200_POOL_GB=<n> in the
 * environment) lets up to that many bytes stay parked per process; 0 switches parking off and frees what is parked.
 * dxmcb200_trim_pool returns the parked blocks of `device` (< 0: all) to the driver. */
int dxmcb200_set_pool_limit(uint64_t bytes);
int dxmcb200_trim_pool(int device);
int dxmcb200_set_luts(dxmcb200_ctx*, const dxmcb200_luts*);
int dxmcb200_set_beam_tables(dxmcb200_ctx*, uint32_t n_spectra, const dxmcb200_spectrum* spectra, uint32_t n_heel,
    const dxmcb200_heel* heel, uint32_t n_bowtie, const dxmcb200_bowtie* bowtie);

/* Scoring is 64-bit fixed point: energy [keV*weight] is added as round(e * 2^energy_bits), energy^2 as
 * round(e^2 * 2^energy_sq_bits) (replaces the float atomics of transport.hpp:208-214, 598-609), so the
 * grids are bit-reproducible for any thread order and any GPU partition. Must be set before the
 * first run and be the same on all ranks; dxmcb200_suggest_fixed_point picks the largest scale that
 * cannot overflow for `total_histories` photons of at most `max_energy_weight` keV*weight each. */
int dxmcb200_suggest_fixed_point(uint64_t total_histories, double max_energy_weight, int* energy_bits, int* energy_sq_bits);
int dxmcb200_set_fixed_point(dxmcb200_ctx*, int energy_bits, int energy_sq_bits);

/* Tracking algorithm of subsequent runs.
 *   0  the reference's Woodcock loop (transport.hpp:640-700) with the global majorant everywhere: draw for draw the
 *      reference's algorithm, what the stream-identical parity tests run;
 *   1  (default) the same loop plus EMPTY-SPACE TRAVERSAL: the grid is cut into bricks of about brick_mm (0 keeps the
 *      current value, default 16 mm; powers of two in voxels, at most 16384 bricks); a brick is "air" when
 *      rho * mu_total(E) <= f_air * majorant(E) for all its voxels and all table energies (f_air <= 0.02) and it holds no
 *      measurement voxel; a photon standing in an air brick (at birth, or after a virtual collision) is walked through the run
 *      of air bricks on its ray (Siddon / Amanatides-Woo traversal of the brick grid) with collision candidates sampled
 *      against f_air * majorant. Delta tracking with any valid majorant samples the same collision density, so results
 *      are statistically equivalent to mode 0 (not draw for draw); the CPU restatement implements the same scheme
 *      (oracle/dxmc_oracle.cpp) and agrees with the kernels stream by stream.
 * Also settable with DXMCB200_TRACKING / DXMCB200_BRICK_MM in the environment at create time. */
int dxmcb200_set_tracking(dxmcb200_ctx*, int tracking, float brick_mm);
/* The brick grid of mode 1 for the uploaded world and tables (built on first use): shift[3] (brick edge = 2^shift voxels),
 * nb[3] bricks per axis, f_air (0: no air bricks); optional (may be NULL) ratio [n_materials] = max_E mu_total,m(E) /
 * majorant(E), brick_max [nb2*nb1*nb0] = max over the brick's voxels of density * ratio[material], air [nb2*nb1*nb0] flags. */
int dxmcb200_get_bricks(dxmcb200_ctx*, uint32_t shift[3], uint32_t nb[3], float* f_air, float* ratio, float* brick_max, uint8_t* air);
/* distance [8][nb2*nb1*nb0]: per octant of travel directions o = (dx<0) | (dy<0)<<1 | (dz<0)<<2 and per brick, the edge k (in
 * bricks, at most 255) of the largest cube of air bricks that has the brick as its corner and opens in the octant's directions
 * (bricks beyond the grid count as air), 0 for non-air bricks: the traversal crosses that cube in one step */
int dxmcb200_get_brick_distance(dxmcb200_ctx*, uint8_t* distance);

/* How the voxel grid is held on the device: 64 = one 8-byte record per voxel, 8 / 4 = palette indices of that many bits per voxel
 * (grids with at most 256 / 16 distinct {density, material, measurement} records; distinct_records, optional, reports how many).
 * With tracking mode 1 the records also say whether the voxel lies in an air brick, which can take a palette grid to the next
 * wider form; the answer is the form the kernels will read. */
int dxmcb200_get_grid_form(dxmcb200_ctx*, int* bits_per_voxel, uint32_t* distinct_records);

/* zero the accumulators and counters */
int dxmcb200_clear(dxmcb200_ctx*);

/* ---- the hot path ------------------------------------------------------------------------- */
/* called on the calling thread after every wave with the number of exposures of the run whose histories have all been
 * issued so far (the value repeats while a long exposure is in flight), and once with the full count at the end; the
 * callee may set the run's cancel flag, which is polled right after the call */
typedef void (*dxmcb200_progress_cb)(uint64_t exposures_done, void* user);

/* Counter-based per-history stream (replaces the per-thread std::random_device seeding of
 * transport.hpp:749): PCG32 (dxmcrandom.hpp:156-165) with {state, increment} =
 * dxmcb200_history_stream(seed, exposure index, history index). */
void dxmcb200_history_stream(uint64_t seed, uint64_t exposure, uint64_t history, uint64_t out_state[2]);

/* Transport exposures[exp_begin, exp_end) of the array `exposures` (indexed absolutely, so any
 * partition of the exposure range over GPUs reproduces the same streams) and add into the ctx's
 * accumulators. low_energy_model: 0 none, 1 Livermore, 2 impulse approximation
 * (lowenergycorrectionmodel.hpp). cancel (may be NULL) is polled between launches. */
int dxmcb200_run(dxmcb200_ctx*, const dxmcb200_exposure* exposures, uint64_t exp_begin, uint64_t exp_end,
    int low_energy_model, uint64_t seed, const volatile int* cancel, dxmcb200_progress_cb cb, void* user);

/* Same with the exposure table already resident on the device: uploaded once by dxmcb200_upload_exposures, or made on the
 * device by dxmcb200_generate_exposures — one kernel evaluates all params->exposures exposures of the source from its
 * parameter block (aec_profile: params->aec_size host floats, may be NULL when aec_size is 0); `out` (may be NULL) receives
 * a host copy of the table. dxmcb200_run_range is dxmcb200_run on the resident table (cancel flag, progress callback). */
int dxmcb200_upload_exposures(dxmcb200_ctx*, const dxmcb200_exposure* exposures, uint64_t n);
int dxmcb200_generate_exposures(dxmcb200_ctx*, const dxmcb200_source_params* params, const float* aec_profile, dxmcb200_exposure* out);
/* device address and length of the resident exposure table (either pointer may be NULL) */
int dxmcb200_exposure_table(dxmcb200_ctx*, void** device_ptr, uint64_t* n);
int dxmcb200_run_range(dxmcb200_ctx*, uint64_t exp_begin, uint64_t exp_end, int low_energy_model, uint64_t seed,
    const volatile int* cancel, dxmcb200_progress_cb cb, void* user);
int dxmcb200_run_resident(dxmcb200_ctx*, uint64_t exp_begin, uint64_t exp_end, int low_energy_model, uint64_t seed);
/* Interleaved partition for multi-GPU runs: transports the resident exposures exp_first + k * exp_stride,
 * k in [0, exp_count). Rank r of N calls (r, N, ceil((n - r) / N)): every GPU then sees the same mix of scan
 * positions ("or strided for load balance"), and because exposures are indexed absolutely the summed
 * grids stay bit-identical to the single-GPU grids. */
int dxmcb200_run_strided(dxmcb200_ctx*, uint64_t exp_first, uint64_t exp_stride, uint64_t exp_count, int low_energy_model,
    uint64_t seed);

/* dxmcb200_run_strided with the cancel flag and progress callback of dxmcb200_run (the callback counts exposures of this rank) */
int dxmcb200_run_strided_monitored(dxmcb200_ctx*, uint64_t exp_first, uint64_t exp_stride, uint64_t exp_count, int low_energy_model,
    uint64_t seed, const volatile int* cancel, dxmcb200_progress_cb cb, void* user);

/* milliseconds the transport kernels of the last run took (CUDA events on the ctx stream) */
int dxmcb200_last_run_ms(dxmcb200_ctx*, double* kernel_ms);

/* ---- outputs ------------------------------------------------------------------------------ */
/* Result<T> with the reference's post-processing (transport.hpp:187-200):
 *   output_mode 0  EV_PER_HISTORY  normalizeScoring (transport.hpp:780-794), n = total_histories
 *   output_mode 1  DOSE            energyImpartedToDose (transport.hpp:796-816) with `calibration`
 *   output_mode 2  raw             dose = sum(e*w) keV, variance = sum((e*w)^2), no scaling
 * Any output pointer may be NULL. Host pointers. */
int dxmcb200_get_result(dxmcb200_ctx*, int output_mode, uint64_t total_histories, float calibration, float* dose,
    uint32_t* n_events, float* variance);
/* raw fixed-point grids (host pointers, any may be NULL) */
int dxmcb200_get_raw(dxmcb200_ctx*, int64_t* energy, uint64_t* energy_sq, uint64_t* events);
/* device address of the accumulator block: n_voxels records of 4 x u64 {energy, energy_sq, events, 0}.
 * Multi-GPU: sum these blocks over ranks (integers: any order gives identical bits) with
 * torch.distributed / NCCL, or call dxmcb200_reduce. */
int dxmcb200_accumulators(dxmcb200_ctx*, void** device_ptr, uint64_t* n_u64);
/* ncclAllReduce(sum, uint64) of the accumulator block on the ctx stream; comm is an ncclComm_t.
 * libnccl is loaded on first use. */
int dxmcb200_reduce(dxmcb200_ctx*, void* nccl_comm);
/* Several GPUs of one process (one ctx and one host thread per device): dxmcb200_comm_create makes the n communicators
 * (ncclCommInitAll) for `devices`; dxmcb200_reduce_collect is called by every rank with its own ctx and communicator after its
 * share of the exposures has run. It sums the fixed-point grids with ONE reduce-scatter over voxel slices (rank r ends up with
 * the total of voxels [r*ceil(V/n), ...)), decodes that slice with the Result post-processing of dxmcb200_get_result and copies
 * it into the caller's host arrays at the slice's offset: all ranks are handed the SAME array pointers and fill disjoint parts
 * (replaces all-reduce of the whole grid + decode on one rank: 1/n of the NVLink traffic per GPU, decode and download in
 * parallel). Integer sums: the result is bit-identical to the single-GPU result. */
int dxmcb200_comm_create(int n, const int* devices, void** nccl_comms);
int dxmcb200_comm_destroy(int n, void** nccl_comms);
int dxmcb200_reduce_collect(dxmcb200_ctx*, void* nccl_comm, int rank, int n_ranks, int output_mode, uint64_t total_histories,
    float calibration, float* dose, uint32_t* n_events, float* variance);

typedef struct dxmcb200_stats {
    uint64_t histories;        /* photons born */
    uint64_t histories_in_world; /* photons that reached the voxel grid */
    uint64_t steps;            /* Woodcock steps taken */
    uint64_t lookups;          /* in-world voxel look-ups (L of the roofline model) */
    uint64_t interactions;     /* real interactions sampled */
    uint64_t score_events;     /* scoring events (S of the roofline model) */
    uint64_t kernel_launches;
    double kernel_ms;          /* summed over all runs since dxmcb200_clear */
    uint64_t air_walks;        /* empty-space traversal: walks through runs of air bricks */
    uint64_t bricks_crossed;   /* ... and brick faces crossed on them */
} dxmcb200_stats;
int dxmcb200_get_stats(dxmcb200_ctx*, dxmcb200_stats*);
/* Device time (CUDA events on the launching streams) and launch counts of the wave kernels since dxmcb200_clear:
 * [0] generateKernel, [1] transportKernel, [2] airWalkKernel (incl. a one-block cursor reset), [3] interactKernel.
 * With two pipelines the kernels of both overlap on the GPU, so the sums can exceed the wall time. */
int dxmcb200_get_kernel_times(dxmcb200_ctx*, double ms[4], uint64_t launches[4]);
/* The per-history work counters (histories .. score_events) cost registers, so the transport kernel is
 * compiled twice; on != 0 selects the counting variant for subsequent runs (default off, or
 * DXMCB200_STATS=1 in the environment at create time). kernel_launches / kernel_ms are always kept. */
int dxmcb200_enable_stats(dxmcb200_ctx*, int on);

/* ---- stand-alone device entry points used by the parity tests ------------------------------
 * Each evaluates one hot-path primitive on the GPU for n inputs (host pointers in/out). */
/* a9/a10: AttenuationLutInterpolator::operator() and maxAttenuationInverse (attenuationinterpolator.hpp:207-248) */
int dxmcb200_eval_attenuation(dxmcb200_ctx*, uint64_t n, const uint8_t* material, const float* energy,
    float* out_photo_compton_rayleigh /* [n][3] */, float* out_max_inverse /* [n] */);
/* Test hook of the empty-space traversal (tracking mode 1): for fixed rays, enter the world like a birth does and, from an air
 * brick, follow the run of air bricks (the Siddon / Amanatides-Woo style traversal over the brick grid, crossing at most
 * DXMCB200_WALK_MAX_CUBES all-air cubes). length [n]: ray parameter of the run's end; info [n]: bits 0-15 cubes crossed, bit 16
 * the ray leaves the grid there, bit 17 the entry point lies in an air brick, bit 18 the ray reaches the world; end [n][3]: the
 * point reached. Bit-identical to the CPU restatement's airRunLength. */
int dxmcb200_trace_air_runs(dxmcb200_ctx*, uint64_t n_rays, const float* pos /*[n][3]*/, const float* dir /*[n][3]*/, float* length, uint32_t* info,
    float* end);

/* a6/a7/a5: for ray r, entry point by transportParticleToWorld then positions pos += dir*step[k], reporting
 * indexFromPosition or -1 once outside (transport.hpp:485-521, 702-728). steps are shared by all rays. */
int dxmcb200_trace_indices(dxmcb200_ctx*, uint64_t n_rays, const float* pos /*[n][3]*/, const float* dir /*[n][3]*/,
    uint32_t n_steps, const float* steps, int64_t* out_indices /*[n][n_steps+1], slot 0 = entry voxel*/,
    float* out_entry /*[n][3]*/);
/* a4: Exposure::sampleParticle for histories [0,n) of exposure `exposure_index` of `exposure`:
 * out [n][8] = pos[3], dir[3], energy, weight (exposure.hpp:280-304) */
int dxmcb200_sample_particles(dxmcb200_ctx*, const dxmcb200_exposure* exposure, uint64_t exposure_index, uint64_t seed,
    uint64_t n, float* out);
/* a13-a16: one interaction of kind (0 photo, 1 Compton, 2 Rayleigh) for n photons of `energy` in
 * `material` travelling along +z, history streams (seed, exposure 0, history i):
 * out [n][5] = energy imparted, new energy, new dir[3] */
int dxmcb200_sample_interaction(dxmcb200_ctx*, int kind, int low_energy_model, uint8_t material, float energy, uint64_t seed,
    uint64_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
