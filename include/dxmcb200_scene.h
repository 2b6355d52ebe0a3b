/* dxmcb200_scene.h — plain-C view of the reference's public C++ API for the transport path.
 *
 * DXMClib has no FFI: its boundary is the C++ call
 *     Result<T> Transport<T>::operator()(const World<T>&, Source<T>*, ProgressBar<T>*, bool)
 * (reference include/dxmc/transport.hpp:135-202) on objects built with the World / Material /
 * Source setters. This header flattens exactly those calls to `extern "C"` so that non-C++
 * hosts (the Python parity tests and bench.py through ctypes; any other FFI) can drive the
 * drop-in C++ classes in dxmclib_b200/include/dxmc/. Each function names the reference member
 * it forwards to. T is float (the reference's validation suite runs float, validation.cpp:1644).
 *
 * The SAME signatures are implemented a second time by oracle/ref_harness.cpp on top of the
 * UNMODIFIED reference headers (library oracle/_ref/libdxmc_ref.so, symbols prefixed the same),
 * which is how the tests put identical scenes through both implementations.
 *
 * All functions return 0 on success, a negative code on error, never throw.
 */
#ifndef DXMCB200_SCENE_H
#define DXMCB200_SCENE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dxs_scene dxs_scene;

enum dxs_status {
    DXS_OK = 0,
    DXS_ERR_ARG = -1,
    DXS_ERR_UNSUPPORTED = -2,
    DXS_ERR_STATE = -3,
    DXS_ERR_DEVICE = -4
};

/* LOWENERGYCORRECTION, reference include/dxmc/lowenergycorrectionmodel.hpp:20-26 */
enum dxs_model { DXS_MODEL_NONE = 0, DXS_MODEL_LIVERMORE = 1, DXS_MODEL_IA = 2 };
/* Transport::OUTPUTMODE, reference transport.hpp:117-120 */
enum dxs_output { DXS_OUT_EV_PER_HISTORY = 0, DXS_OUT_DOSE = 1 };

/* which implementation this library is: "dxmc-b200" or "dxmclib-reference" */
const char* dxs_backend(void);
/* message of the last failed call on this thread ("" when none) */
const char* dxs_last_error(void);

dxs_scene* dxs_create(void);
void dxs_destroy(dxs_scene*);

/* ---- World<T> (reference world.hpp:37-227) ------------------------------------------- */
/* setDimensions / setSpacing / setOrigin / setDirectionCosines */
int dxs_world_geometry(dxs_scene*, const uint64_t dim[3], const float spacing[3], const float origin[3],
    const float cosines[6]);
/* addMaterialToMap(Material(nameOrFormula, "", density)); density<=0 keeps the material's own */
int dxs_world_add_material(dxs_scene*, const char* name_or_formula, double density);
/* addMaterialToMap(Material(Z)) */
int dxs_world_add_element(dxs_scene*, int Z);
/* setDensityArray / setMaterialIndexArray / setMeasurementMapArray (measurement may be NULL); copied */
int dxs_world_arrays(dxs_scene*, const float* density, const uint8_t* material, const uint8_t* measurement);
/* replace the world by CTDIPhantom<T>(diameter) (reference world.hpp:229-361) */
int dxs_world_ctdi_phantom(dxs_scene*, uint64_t diameter_mm);
/* makeValid(); returns 1 valid / 0 invalid in *valid */
int dxs_world_validate(dxs_scene*, int* valid);
int dxs_world_dimensions(dxs_scene*, uint64_t dim[3], float spacing[3], float extent_safe[6]);
/* copies of the voxel arrays (any pointer may be NULL) */
int dxs_world_get_arrays(dxs_scene*, float* density, uint8_t* material, uint8_t* measurement);
/* CTDIPhantom::holeIndices(pos) pos: 0 centre,1 west,2 east,3 south,4 north; call with out==NULL for the count */
int dxs_world_ctdi_holes(dxs_scene*, int position, uint64_t* out, uint64_t* count);

/* Voxel-index sequence of fixed rays through the world: transportParticleToWorld, then pos += dir*step[k] with
 * particleInsideWorld / indexFromPosition after every step (reference transport.hpp:485-521, 702-728). out_indices is
 * [n_rays][n_steps+1] (slot 0 = voxel of the entry point, -1 = outside), out_entry [n_rays][3]. The reference harness
 * calls the reference's own members; the product evaluates the CUDA device functions (needs a GPU). */
int dxs_trace_indices(dxs_scene*, uint64_t n_rays, const float* pos, const float* dir, uint32_t n_steps, const float* steps,
    int64_t* out_indices, float* out_entry);

/* ---- Material (reference material.hpp:62-104) evaluated for world material `idx` ----- */
int dxs_material_attenuation(dxs_scene*, int idx, double energy, double out_photo_compton_rayleigh_total[4]);
int dxs_material_form_factor_sq(dxs_scene*, int idx, double q, double* out);
int dxs_material_scatter_factor(dxs_scene*, int idx, double q, double* out);
/* getBindingEnergies(minValue); call with out==NULL for the count */
int dxs_material_binding_energies(dxs_scene*, int idx, double min_value, double* out, int* count);
/* getElectronConfiguration(): 12 shells x 13 doubles
 * {binding, nElectrons, J0, photoIonProb, yield, lineProb[3], lineEnergy[3], Z, shell} */
int dxs_material_shells(dxs_scene*, int idx, double out[12 * 13]);
int dxs_material_density(dxs_scene*, int idx, double* out);

/* ---- AttenuationLut<T> (reference attenuationlut.hpp:42-275) ------------------------- */
/* generate(world, maxEnergy, minEnergy=1) */
int dxs_lut_generate(dxs_scene*, float max_energy);
int dxs_lut_attenuation(dxs_scene*, int material, float energy, float out[3]);      /* photoComptRayAttenuation */
int dxs_lut_max_inverse(dxs_scene*, float energy, float* out);                      /* maxTotalAttenuationInverse */
int dxs_lut_scatter_factor(dxs_scene*, int material, float q, float* out);          /* comptonScatterFactor */
/* momentumTransferFromFormFactor driven by a seeded PCG32 RandomState; n samples */
int dxs_lut_sample_form_factor(dxs_scene*, int material, float qmax_squared, uint64_t seed[2], int n, float* out);
/* raw tables for bit-exact comparison. sizes are returned when the pointer is NULL.
 * what: 0 knots m_x, 1 coefficients, 2 majorant coefficients,
 *       3 {linearIndex, linearStep, linearEnergy, resolution} as 4 floats,
 *       4 RITA x|e|a|b (4*56 per material), 5 spline coeffs+x+step/start/stop (60+16+3 per material)
 *       16 + k (product library only): table k of the Transport prepared by dxs_b200_prepare */
int dxs_lut_table(dxs_scene*, int what, float* out, uint64_t* count);

/* ---- Sources (reference source.hpp) ---------------------------------------------------- */
/* PencilSource: setPosition, setDirectionCosines, setPhotonEnergy, setHistoriesPerExposure, setTotalExposures */
int dxs_source_pencil(dxs_scene*, const float pos[3], const float cosines[6], float energy, uint64_t histories,
    uint64_t exposures);
/* IsotropicSource / IsotropicCTSource (ct!=0): setSpecter(weights, energies), setCollimationAngles(x0,x1,y0,y1) */
int dxs_source_isotropic(dxs_scene*, int ct, const float pos[3], const float cosines[6], const float collimation[4],
    int n_spectrum, const float* weights, const float* energies, uint64_t histories, uint64_t exposures);

typedef struct dxs_tube {
    float voltage;           /* Tube::setVoltage */
    float anode_angle_deg;   /* setAnodeAngleDeg */
    float al_mm, cu_mm, sn_mm; /* set{Al,Cu,Sn}Filtration; <=0 leaves unset */
    float energy_resolution; /* setEnergyResolution; <=0 keeps 1 keV */
} dxs_tube;

/* DXSource: tube, setSourceDetectorDistance, setFieldSize, setSourceAnglesDeg, setTubeRotationDeg, setDap */
typedef struct dxs_dx_params {
    dxs_tube tube;
    float position[3];
    float sdd;
    float field_size[2];
    float source_angles_deg[2];
    float tube_rotation_deg;
    float dap;
    int model_heel;
    uint64_t histories, exposures;
} dxs_dx_params;
int dxs_source_dx(dxs_scene*, const dxs_dx_params*);

/* CTAxialSource (spiral==0) / CTSpiralSource (spiral!=0) */
typedef struct dxs_ct_params {
    dxs_tube tube;
    int spiral;
    float position[3];
    float cosines[6];        /* all zero keeps the CT default {-1,0,0,0,0,1} */
    float sdd, collimation, fov;
    float start_angle_deg, exposure_step_deg;
    float scan_length;
    float pitch;             /* spiral */
    float step;              /* axial; <=0 keeps collimation */
    float gantry_tilt_deg;
    float ctdi_vol;
    uint64_t ctdi_phantom_diameter;
    int model_heel;
    int use_xcare;
    float xcare_filter_angle_deg, xcare_span_deg, xcare_ramp_deg, xcare_low_weight;
    uint64_t histories;
} dxs_ct_params;
int dxs_source_ct(dxs_scene*, const dxs_ct_params*);
/* CTAxialDualSource / CTSpiralDualSource: tube A and the scan as in dxs_ct_params, plus tube B (setSourceDetectorDistanceB,
 * setFieldOfViewB, setStartAngleDegB, tubeB(), setTubeAmas / setTubeBmas); <= 0 keeps the defaults */
typedef struct dxs_ct_dual_params {
    dxs_ct_params a;
    dxs_tube tube_b;
    float sdd_b, fov_b, start_angle_b_deg, mas_a, mas_b;
} dxs_ct_dual_params;
int dxs_source_ct_dual(dxs_scene*, const dxs_ct_dual_params*);
/* CTTopogramSource from the tube and geometry fields of dxs_ct_params (position, cosines, sdd, collimation, fov, start angle,
 * gantry tilt, scan_length, histories) */
int dxs_source_topogram(dxs_scene*, const dxs_ct_params*);
/* CBCTSource: the DX fields (exposures ignored) plus setSpanAngleDeg / setStepAngleDeg */
typedef struct dxs_cbct_params {
    dxs_dx_params dx;
    float span_deg, step_deg;
} dxs_cbct_params;
int dxs_source_cbct(dxs_scene*, const dxs_cbct_params*);
/* setBowTieFilter(BowTieFilter(angles, weights)) on a CT source */
int dxs_source_bowtie(dxs_scene*, int n, const float* angles_rad, const float* weights);
/* setAecFilter(AECFilter(world density, spacing, dims, exposure profile along z[nz])) on a CT source */
int dxs_source_aec(dxs_scene*, int n, const float* exposure_profile);

int dxs_source_total_exposures(dxs_scene*, uint64_t* n);
int dxs_source_max_energy(dxs_scene*, float* e);

/* One exposure after source->validate(), as the transport workers see it
 * (getExposure(i) then alignToDirectionCosines(world basis), reference transport.hpp:756-757). */
typedef struct dxs_exposure {
    float position[3];
    float cosines[6];
    float beam_direction[3];
    float collimation[4];
    float weight;
    float mono_energy;      /* m_monoenergeticPhotonEnergy */
    int32_t has_spectrum;   /* m_specterDistribution != nullptr */
    int32_t has_heel;       /* m_heelFilter != nullptr */
    int32_t has_bowtie;     /* m_beamFilter != nullptr */
    uint64_t histories;
} dxs_exposure;
int dxs_source_exposure(dxs_scene*, uint64_t i, dxs_exposure* out);
/* The sampling tables exposure 0 points to, for bit-exact comparison and for building dxmcb200_* inputs by hand.
 * what: 0 spectrum m_probs, 1 spectrum m_alias (as floats), 2 spectrum m_energies   (dxmcrandom.hpp:277-329)
 *       3 heel {energyStart, energyStep, energySize, angleStart, angleStep, angleSize}, 4 heel m_weights (beamfilters.hpp:438-448)
 *       5 bow-tie |angle|, 6 bow-tie normalised weight                                 (beamfilters.hpp:80, 199)
 * A missing table has count 0. out==NULL returns the count only. */
int dxs_source_table(dxs_scene*, int what, float* out, uint64_t* count);
/* the source's normalised spectrum after validate() (tube model or user spectrum); NULL out for count */
int dxs_source_spectrum(dxs_scene*, float* energies, float* weights, int* count);
/* Source::getCalibrationValue(model) */
int dxs_source_calibration(dxs_scene*, int model, float* out);

/* ---- Transport<T>::operator() ------------------------------------------------------------
 * seed: the reference seeds each worker from std::random_device (transport.hpp:749), so a stock run is
 * not reproducible; with seed!=0 the reference harness runs the workers' loop (getExposure, align,
 * transport<L>) on ONE thread with RandomState{seed, seed^0x9E3779B97F4A7C15}; seed==0 runs the
 * stock multi-threaded operator(). The B200 implementation keys its per-history counter streams on
 * `seed` in both cases. n_workers==0: hardware_concurrency (reference default).
 * n_workers==DXS_WORKERS_COUNTER_STREAMS: the reference harness gives EVERY history its own RandomState
 * seeded by the product's dxmcb200_history_stream(seed, exposure, history) and spreads exposures over all
 * host threads, so both implementations consume identical random numbers history by history and
 * differ only by libm-vs-CUDA rounding. */
#define DXS_WORKERS_COUNTER_STREAMS (-1)
typedef struct dxs_result_info {
    uint64_t histories;
    double seconds;          /* Result::simulationTime */
    char units[16];          /* Result::dose_units */
} dxs_result_info;
int dxs_transport(dxs_scene*, int model, int output_mode, int use_calibration, uint64_t seed, int n_workers,
    float* dose, uint32_t* n_events, float* variance, dxs_result_info* info);

/* Transport::operator() WITH a ProgressBar, driven the way the reference's callers drive it (validation.cpp:269-290):
 * the call runs on this thread while a second thread polls ProgressBar::getETA and computeDoseProgressImage every 2 ms
 * and, when cancel_at_percent > 0, calls setCancel(true) once the reported progress reaches it. A cancelled run returns
 * an all-zero result with histories == 0 (transport.hpp:176-185). Both implementations export it. */
typedef struct dxs_progress_report {
    double percent_seen;      /* largest progress the monitor saw in getETA() while the run was in flight */
    double percent_final;     /* progress in getETA() after the call returned */
    uint32_t images_polled;   /* computeDoseProgressImage() calls that returned an image */
    uint32_t images_nonzero;  /* ... of the right size with at least one non-zero pixel (the live dose is visible) */
    uint32_t image_width, image_height;
    int32_t cancelled;        /* the monitor called setCancel(true) */
    char eta[96];             /* last getETA() string */
} dxs_progress_report;
int dxs_transport_monitored(dxs_scene*, int model, int output_mode, int use_calibration, uint64_t seed, int n_workers,
    double cancel_at_percent, float* dose, uint32_t* n_events, float* variance, dxs_result_info* info,
    dxs_progress_report* report);

/* ---- B200 extensions: the phases of Transport::operator() kept apart so a caller can hold the world,
 * tables and exposures resident on the GPU (bench.py device-resident timing, multi-GPU sharding).
 * The reference harness answers DXS_ERR_UNSUPPORTED. -------------------------------------------------- */
/* Transport::prepare: validate, build + upload LUTs, world, beam tables and ALL exposures on `device`.
 * total_histories_all_ranks sizes the fixed-point scale for a run sharded over several GPUs (0: this scene). */
int dxs_b200_prepare(dxs_scene*, int device, int model, uint64_t seed, uint64_t total_histories_all_ranks);
/* Transport::run on exposures [begin, end); kernel_ms (may be NULL) = CUDA-event time of the transport kernels */
int dxs_b200_run(dxs_scene*, uint64_t exp_begin, uint64_t exp_end, double* kernel_ms);
/* Transport::runStrided: exposures exp_first + k * exp_stride, k < exp_count (interleaved multi-GPU partition) */
int dxs_b200_run_strided(dxs_scene*, uint64_t exp_first, uint64_t exp_stride, uint64_t exp_count, double* kernel_ms);
/* Transport::collect: decode accumulators into Result arrays (any pointer may be NULL) */
int dxs_b200_collect(dxs_scene*, int output_mode, int use_calibration, uint64_t histories, float* dose, uint32_t* n_events,
    float* variance, dxs_result_info* info);
/* the prepared dxmcb200_ctx* (include/dxmcb200.h) for direct C-ABI calls: accumulators, stats, clear */
int dxs_b200_context(dxs_scene*, void** ctx);
int dxs_b200_release(dxs_scene*);
/* Transport::setDevices for subsequent dxs_transport / dxs_transport_monitored calls: with more than one device the one call
 * spreads the exposures over all of them (one host thread per GPU, NCCL reduce-scatter of the dose grids). n == 0: back to one GPU. */
int dxs_b200_set_devices(dxs_scene*, int n, const int* devices);

#ifdef __cplusplus
}
#endif
#endif
