/*
 * lisflood_b200.h -- C ABI of liblisf_b200.so, the B200 (sm_100a) implementation of LISFLOOD's
 * raster time-step hot path (kinematic-wave routing + per-cell soil / surface-runoff updates).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy types.  The Python host
 * layer (lisflood_code_b200/) binds it with ctypes and mirrors the reference's operator API on top;
 * INTEGRATION.md shows the stub a LISFLOOD maintainer would add.
 *
 * Conventions
 *   - every function returns LF_OK (0) or a negative LF_ERR_* code; lf_last_error() returns a
 *     thread-local description of the last failure.  No C++ exception crosses the ABI.
 *   - opaque handles own device memory; host pointers are borrowed for the duration of a call.
 *   - thread-compatible, not thread-safe (the reference drives everything from one Python thread).
 *   - array arguments may be HOST pointers or DEVICE pointers of the same CUDA context (unified virtual
 *     addressing; the library copies with cudaMemcpyDefault) unless stated otherwise.
 *   - all floating point is IEEE float64; "compressed" arrays are the reference's 1-D arrays of
 *     active pixels in row-major order of the mask (global_modules/add1.py:268-282).
 *   - reference citations are relative to /root/reference/src/lisflood/.
 */
#ifndef LISFLOOD_B200_H
#define LISFLOOD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LF_OK 0
#define LF_ERR_INVALID (-1)   /* bad argument (null pointer, size, section ...) */
#define LF_ERR_BAD_LDD (-2)   /* an LDD code outside 0..9 */
#define LF_ERR_LDD_CYCLE (-3) /* the LDD contains a cycle (the reference would loop forever,
                                 hydrological_modules/kinematic_wave_parallel.py:99) */
#define LF_ERR_CUDA (-4)      /* CUDA runtime failure, text in lf_last_error() */
#define LF_ERR_NO_DEVICE (-5) /* no sm_100 device: there is NO CPU fallback */
#define LF_ERR_STATE (-6)     /* call order violated (e.g. floodplain section without alpha_fp) */

#define LF_SECTION_MAIN 0       /* section="main_channel" */
#define LF_SECTION_FLOODPLAIN 1 /* section="floodplains"  */

typedef struct lf_graph lf_graph;   /* D8 drainage graph + routing order, device resident */
typedef struct lf_router lf_router; /* one kinematicWave object */
typedef struct lf_model lf_model;   /* full hot-path step: soil -> overland -> channel sub-steps */
typedef struct lf_xchg lf_xchg;     /* exchange region of one rank (LDD-cut decomposition across the GPUs of a node) */

const char *lf_last_error(void);
int lf_version(void);

/* Selects the CUDA device (default 0) and verifies it is compute capability 10.x.
 * Every other entry point fails with LF_ERR_NO_DEVICE when there is no such device. */
int lf_device_init(int device);
/* name: >= 128 bytes.  Any out pointer may be NULL. */
int lf_device_info(char *name, int *sm_count, int *cc_major, int *cc_minor, int64_t *hbm_bytes);
/* Blocks until all work queued by this library on its stream has finished. */
int lf_synchronize(void);
/* Device-side stopwatch on the library's stream (CUDA events). */
int lf_timer_start(void);
int lf_timer_stop(double *elapsed_ms);
/* Counts kernels launched by this library since the last reset (bench.py's gpu_launches). */
int64_t lf_launch_count(int reset);
/* Launch API calls the host made since that reset: a replayed CUDA graph counts once here and with all its kernels in
 * lf_launch_count. */
int64_t lf_host_launch_count(void);
/* Page-locks / unlocks a host buffer the caller will pass repeatedly (faster H2D/D2H). */
int lf_host_register(void *ptr, int64_t bytes);
int lf_host_unregister(void *ptr);
/* Page-locked host memory owned by the library (for buffers whose lifetime the caller cannot tie to an unregister call:
 * a range that is freed while still registered poisons later copies from re-used addresses). */
int lf_host_alloc(int64_t bytes, void **out);
int lf_host_free(void *ptr);
/* On-device accuracy check of the library's hand-written float64 math against the CUDA math library over n
 * pseudo-random arguments: max_err[8] = maximum relative error of { Newton division, Newton square root, table x^y
 * (normalised by 1 + |y log2 x|), table e^x (normalised by 1 + |x|), fifth root, cube root, polynomial x^y
 * (normalised likewise), van Genuchten term 1 - (1 - s^(1/m))^m }.  Test infrastructure (tests/test_gpu_math.py). */
int lf_math_selftest(int64_t n, uint64_t seed, double *max_err);

/* ---------------------------------------------------------------------------------------------
 * Drainage graph.  Replaces kinematicWave.__init__'s graph part:
 *   rebuildFlowMatrix/decodeFlowMatrix  hydrological_modules/kinematic_wave_parallel.py:59-71
 *   streamLookups + upDownLookups       :73-90, kinematic_wave_parallel_tools.py:111-130
 *   topoDistFromSea + _setRoutingOrders :92-106, :140-158
 * ldd_codes: f64[N] compressed keypad codes (1-9, 5 and 0 = pit); land_mask: u8[rows*cols], 1 =
 * active pixel, N = number of ones.  The exported arrays are bit-identical to the reference's.
 * ------------------------------------------------------------------------------------------- */
int lf_ldd_build(const double *ldd_codes, const uint8_t *land_mask, int64_t rows, int64_t cols, lf_graph **out);
int lf_graph_info(const lf_graph *g, int64_t *n_pixels, int64_t *n_orders, int64_t *max_upstream,
                  int64_t *n_pits);
/* Any pointer may be NULL.  pixels_ordered i64[N]; order_start_stop i64[n_orders*2];
 * upstream_lookup i64[N*max_upstream] (-1 fill); num_upstream i64[N]; downstream f64[N] (-1 = none). */
int lf_graph_export(const lf_graph *g, int64_t *pixels_ordered, int64_t *order_start_stop,
                    int64_t *upstream_lookup, int64_t *num_upstream, double *downstream);
/* Internal storage order ("position" = breadth-first layout from the outlets):
 * pixel_of_position i32[N], level_start i32[n_orders+1].  For tests / diagnostics. */
int lf_graph_layout(const lf_graph *g, int32_t *pixel_of_position, int32_t *level_start);
/* PCRaster accuflux(ldd, x): downstream-accumulated sum of x including the cell itself, f64[N] compressed
 * (init-time operator of the reference, hydrological_modules/routing.py:98). */
int lf_graph_accuflux(const lf_graph *g, const double *x, double *out);
void lf_graph_destroy(lf_graph *g);

/* ---------------------------------------------------------------------------------------------
 * Kinematic-wave router.  Replaces class kinematicWave
 * (hydrological_modules/kinematic_wave_parallel.py:114-184) and the Numba kernels
 * kinematicRouting / solve1Pixel / closureError (kinematic_wave_parallel_tools.py:34-92).
 * alpha f64[N]; dx f64[N] or NULL (then dx_scalar); alpha_floodplains f64[N] or NULL.
 * The graph must outlive the router.
 * ------------------------------------------------------------------------------------------- */
int lf_router_create(lf_graph *g, const double *alpha, double beta, const double *dx, double dx_scalar,
                     double dt, const double *alpha_floodplains, int flagnancheck, lf_router **out);
/* kinematicWaveRouting(discharge, specific_lateral_inflow, section): host f64[N] in compressed
 * order; discharge is updated IN PLACE (H2D, solve, D2H inside the call).
 * nonfinite (may be NULL) receives 1 when flagnancheck is set and a NaN/Inf was produced
 * (the reference warns once, it does not fail: kinematic_wave_parallel.py:180-184). */
int lf_router_route(lf_router *r, double *discharge, const double *specific_lateral_inflow, int section,
                    int *nonfinite);

/* Device-resident variant: state stays in HBM between calls. */
int lf_router_set_discharge(lf_router *r, int section, const double *discharge);
int lf_router_get_discharge(lf_router *r, int section, double *discharge);
int lf_router_set_inflow(lf_router *r, int section, const double *specific_lateral_inflow);
/* Runs nsteps consecutive kinematicWaveRouting calls on the resident state as ONE space-time
 * wavefront: step s uses lateral inflow q * q_scale[s] (q_scale NULL = all ones;
 * host f64[nsteps]).  Same arithmetic per (pixel, step) as nsteps calls of lf_router_route. */
int lf_router_run(lf_router *r, int section, int nsteps, const double *q_scale, int *nonfinite);
/* LDD-cut domain decomposition: see the "Multi-GPU" block below (lf_xchg_*, lf_router_set_exchange). */
/* Execution options of a router: "cuda_graphs" 1 (default): the diagonals of a run are captured once and replayed as a
 * CUDA graph; "cooperative" k > 0: a run over a deep network is ONE persistent cooperative launch (k resident blocks per
 * SM) with a grid barrier per diagonal -- measured slower than graph replay on B200, off by default; both 0: one plain
 * launch per diagonal.  "narrow_runs" 1: consecutive diagonals of at most 512 work items run back to back in ONE
 * single-block launch, a block barrier between two diagonals; 0 (default; measured: no gain): one launch each. */
int lf_router_set_option(lf_router *r, const char *name, double value);
void lf_router_destroy(lf_router *r);

/* ---------------------------------------------------------------------------------------------
 * Full hot-path step on device-resident state.  One lf_model holds what the reference keeps on the
 * shared model object `self.var` for these modules (attribute names: SURVEY.md A.3) and replaces, per
 * model time step (Lisflood_dynamic.py:114-229):
 *   lf_model_soil            soilloop.dynamic_canopy + dynamic_soil   hydrological_modules/soilloop.py:519-665
 *                            opensealed.dynamic                        hydrological_modules/opensealed.py:41-71
 *                            soil.dynamic_perpixel                     hydrological_modules/soil.py:471-514
 *                            groundwater.dynamic                       hydrological_modules/groundwater.py:134-180
 *                            runoff components of surface_routing      hydrological_modules/surface_routing.py:122-149
 *   lf_model_surface_routing surface_routing.dynamic (3 routers)       hydrological_modules/surface_routing.py:151-212
 *   lf_model_channel         NoRoutSteps x routing.dynamic + post-loop hydrological_modules/routing.py:435-706,
 *                                                                       Lisflood_dynamic.py:176-229
 *   lf_model_step            the three stages in order.
 * Options covered: kinematic wave (no dynamicWave), single or split routing, no structures / water use /
 * inflow / transmission loss / open-water evaporation (out of scope).
 * ------------------------------------------------------------------------------------------- */
typedef struct lf_model_config {
    int64_t rows, cols;
    double DtSec;            /* model time step [s]; DtDay = DtSec / 86400 */
    double Beta;             /* kinematic wave beta */
    double PixelLength;      /* [m] (scalar, like the reference without gridSizeUserDefined) */
    int32_t NoRoutSteps;     /* channel sub-steps per model step; DtRouting = DtSec / NoRoutSteps */
    int32_t SplitRouting;    /* 0 single kinematic routing, 1 main channel + floodplain */
    double CourantCrit, AvWaterThreshold, LeafDrainageK, DrainedFraction, SMaxSealed;
    int32_t diagnostics;     /* 1: materialise every flux / diagnostic map of the reference (tests, reporting) */
    int32_t reserved;
} lf_model_config;

/* land_mask u8[rows*cols]; ldd_to_chan, ldd_kinematic: f64[N] compressed keypad codes of LddToChan and
 * LddKinematic (hydrological_modules/routing.py:118-153). */
int lf_model_create(const lf_model_config *cfg, const uint8_t *land_mask, const double *ldd_to_chan,
                    const double *ldd_kinematic, lf_model **out);
int lf_model_info(const lf_model *m, int64_t *n_pixels, int64_t *levels_overland, int64_t *levels_channel,
                  int64_t *isolated_channel_pixels, int64_t *device_bytes);
/* Named maps in the reference's compressed order: count = N for per-pixel maps, 3*N for
 * (vegetation|landuse|runoff, pixel) maps.  Unknown names -> LF_ERR_INVALID. */
int lf_model_set(lf_model *m, const char *name, const double *values, int64_t count);
/* Like lf_model_set for per-pixel / (vegetation, pixel) maps, but returns immediately: the host-to-device copy
 * runs on a separate copy stream (so it overlaps the kernels of the step in flight) and the map takes its new
 * value after all work already queued.  `values` must stay valid and unchanged until the next synchronising call
 * (lf_model_get, lf_synchronize); use page-locked memory for a truly asynchronous copy. */
int lf_model_set_async(lf_model *m, const char *name, const double *values, int64_t count);
int lf_model_get(lf_model *m, const char *name, double *values, int64_t count);
/* Like lf_model_get, but returns at once: the map is brought to the reference's order on the compute stream (after all
 * work already queued) and copied to `values` on a separate output stream, so the copy and whatever the host does with
 * the previous step's map (writing dis.nc, time series) overlap the next step's kernels.  `values`
 * (page-locked for a real overlap) is valid after lf_model_wait_outputs. */
int lf_model_get_async(lf_model *m, const char *name, double *values, int64_t count);
/* The same, narrowed to float32 on the device first: OutputMapsDataType = float32 of the reference's map writer
 * (global_modules/netcdf.py:478), half the bytes over the host link. */
int lf_model_get_async_f32(lf_model *m, const char *name, float *values, int64_t count);
int lf_model_wait_outputs(lf_model *m);
/* boolean maps (u8[N]): "isFrozenSoil", "IsChannel", "IsChannelKinematic", "AtLastPointC" */
int lf_model_set_flags(lf_model *m, const char *name, const uint8_t *values, int64_t count);
int lf_model_soil(lf_model *m);
int lf_model_surface_routing(lf_model *m);
int lf_model_channel(lf_model *m);
int lf_model_step(lf_model *m);
/* Device time (CUDA events on the library stream) spent in the three stages of lf_model_step since the last
 * reset, in milliseconds, and the number of steps they cover.  Synchronises. */
int lf_model_stage_times(lf_model *m, int reset, double *soil_ms, double *overland_ms, double *channel_ms,
                         int64_t *steps);
/* Soil-stage statistics of the last step: deferred_columns[6] = number of (fraction, pixel) columns that needed
 * 2-3, 4-7, 8-15, 16-31, 32-63, 64+ Darcy sub-steps; kernel_ms[8] = device time of the first pass (k_soil_staged;
 * k_soil_fused with diagnostics) in [0], of k_soil_veg_deferred (one persistent launch over all six lists) in [1],
 * zeros in [2..6], of k_soil_pixel_flagged in [7] (all zeros unless timing was enabled by an earlier call with
 * enable_timing = 1).  Either pointer may be NULL.  Synchronises. */
int lf_model_soil_stats(lf_model *m, int enable_timing, int64_t *deferred_columns, double *kernel_ms);
/* Feeder modules of a step on the device: the scaling of readmeteo.dynamic
 * (hydrological_modules/readmeteo.py:61-81), snow.dynamic (snow.py:95-187: three elevation zones, seasonal melt
 * coefficient, summer ice melt) and frost.dynamic (frost.py:61-78), fused in one kernel that takes the RAW meteo maps of
 * the step -- Precipitation [mm/day], Tavg [deg C], ET0, E0 [mm/day]; compressed order; dtype 0 = float64, 1 = float32
 * (what the NetCDF forcing holds; widened to float64 exactly, all arithmetic is float64) -- and leaves Rain, SnowMelt,
 * ETRef, EWRef, ESRef, isFrozenSoil for the soil stage, updates the states SnowCoverS (3,N), FrostIndex,
 * TotalPrecipitation and (diagnostics build) writes Snow, SnowCover, Precipitation, Tavg.  The three season coefficients
 * are the scalars of snow.py:104-117 for the calendar day (computed by the host mirror exactly as the reference does).
 * Parameters by the reference's names -- PrScaling, CalEvaporation, DeltaTSnow, SnowSeason, TempSnow, SnowFactor,
 * SnowMeltCoef, TempMelt, lat_rad, Kfrost, Afrost, FrostIndexThreshold, SnowWaterEquivalent, kgb -- are maps
 * (lf_model_set) or scalars (lf_model_set_scalar), like the float-or-array results of the reference's loadmap.
 * async = 1: host buffers (page-locked for a real overlap) are copied on the copy stream while the previous step computes
 * and must stay untouched until the next synchronising call; two staging sets alternate.
 * lf_model_set_lai: LAI (3,N) of the current 10-day interval and LAITerm = exp(-kgb * LAI) (leafarea.py:48,90). */
int lf_model_set_scalar(lf_model *m, const char *name, double value);
int lf_model_feed(lf_model *m, const void *precipitation, const void *tavg, const void *et0, const void *e0, int32_t dtype,
                  double snowmelt_coeff, double ice_melt_coeff_north, double ice_melt_coeff_south, int32_t async);
/* The same for forcing still packed the CF way -- int16 with the variable's scale_factor / add_offset attributes, which the
 * reference's reader applies on the host (global_modules/netcdf.py:231-232, xarray's CF decoding): value = raw * scale_factor
 * + add_offset, here on the device, so that half the bytes of the float32 maps cross the host link.  scale_factor /
 * add_offset: 4 values each (Precipitation, Tavg, ET0, E0).  decode_float32 = 1 unpacks in float32 arithmetic (what a CF
 * decoder yields for int16 data with float32 attributes), 0 in float64.  Missing-value codes are not looked at: compressed
 * maps hold catchment pixels only. */
int lf_model_feed_packed(lf_model *m, const int16_t *precipitation, const int16_t *tavg, const int16_t *et0, const int16_t *e0,
                         const double *scale_factor, const double *add_offset, int32_t decode_float32, double snowmelt_coeff,
                         double ice_melt_coeff_north, double ice_melt_coeff_south, int32_t async);
int lf_model_set_lai(lf_model *m, const double *lai, int64_t count);
/* Structures inside the routing sub-step loop: reservoirs (four-regime outflow rule,
 * hydrological_modules/reservoir.py:173-322) and lakes (Modified Puls, lakes.py:199-297), executed by the channel
 * wavefront itself.  The model must have been created with ldd_kinematic = the channel network BEFORE
 * structures.initial cuts it (`LddStructuresKinematic`, structures.py:43-61): the library applies the cut (nothing is
 * routed into a structure pixel; its inflow is the ChanQ its upstream neighbours had after the previous sub-step,
 * np.bincount(downstruct, ChanQ), reservoir.py:190 / lakes.py:215; its outflow joins the side flow, routing.py:472-476).
 * reservoir_index / lake_index: compressed pixel indices (ReservoirIndex, LakeIndex), host arrays.  On a cut raster
 * every rank passes the structures it owns (local indices); a structure and the pixels draining into it must sit on one
 * rank (lisflood_code_b200/parallel.py arranges that).
 * lf_model_structure_array reads (set = 0) or writes (set = 1) a per-structure array (host or device pointer), in the
 * order of the index arrays, by the reference's names: TotalReservoirStorageM3CC, ConservativeStorageLimitCC,
 * NormalStorageLimitCC, Normal_FloodStorageLimitCC, FloodStorageLimitCC, MinReservoirOutflowCC, NormalReservoirOutflowCC,
 * NonDamagingReservoirOutflowCC, DeltaO, DeltaLN, DeltaNFL, ReservoirStorageM3CC (state), ReservoirFillCC, QResOutM3DtCC
 * (outputs of the last sub-step); LakeAreaCC, LakeFactor, LakeFactorSqr, LakeStorageM3CC, LakeOutflowCC, LakeInflowOldCC,
 * LakeStorageM3BalanceCC (state), LakeLevelCC, QLakeOutM3DtCC (outputs). */
int lf_model_set_structures(lf_model *m, int32_t n_reservoirs, const int64_t *reservoir_index, int32_t n_lakes,
                            const int64_t *lake_index);
int lf_model_structure_array(lf_model *m, const char *name, double *values, int64_t count, int32_t set);
/* Execution options of a model (name, value):
 *   "overlap_isolated"     1: lf_model_step starts the sub-steps of the non-channel isolated pixels of
 *                          LddKinematic (no side flow, routing.py:512) at the top of the step, on a low-priority
 *                          stream alongside the soil stage; 0 (default; the step is bound by the FP64 pipe, co-running
 *                          the two did not shorten it): everything of the channel stage runs inside it.
 *   "early_blocks_per_sm"  resident blocks per SM of that early launch (default 2).
 *   "isolated_blocks_per_sm" resident blocks per SM of the isolated-pixel kernel inside the channel stage (default 6:
 *                          leaves room for the wavefront's blocks, so the two overlap; 0: one block per chunk).
 *   "narrow_runs"          1: consecutive wavefront diagonals of at most 512 work items run back to back in ONE
 *                          single-block launch (a block barrier between two diagonals instead of a kernel boundary);
 *                          0 (default; measured: no gain): one launch per diagonal.
 *   "cuda_graphs"          1 (default): the level sweep of the overland routers and the diagonals of the channel
 *                          wavefront (hundreds of small dependent launches per step) are captured once per argument
 *                          set and replayed as CUDA graphs; 0: plain launches.
 *   "accumulate_discharge" 1: CumQ += ChanQ after every step (options InitLisflood / repAverageDis,
 *                          Lisflood_dynamic.py:224-227); avgdis = CumQ / steps is the pre-run product `AvgDis`.
 *   "flagnancheck"         1: the `-n` option of the reference (kinematic_wave_parallel.py:180-184) for the model's
 *                          channel discharge: a non-finite ChanQ raises a flag read by lf_model_nonfinite. */
int lf_model_set_option(lf_model *m, const char *name, double value);
/* *nonfinite = 1 when a NaN/Inf channel discharge was produced since the last call (needs "flagnancheck");
 * the reference warns once and goes on (kinematic_wave_parallel.py:180-184).  Synchronises. */
int lf_model_nonfinite(lf_model *m, int *nonfinite);
void lf_model_destroy(lf_model *m);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: ONE catchment raster cut along its drainage graph over the GPUs of a node, one process per GPU
 * (the reference shows that sub-catchment runs reproduce the full run bit for bit,
 * tests/test_subcatchments.py:111-112).  Host side: lisflood_code_b200/parallel.py.
 *
 *   lf_graph_partition   owner rank of every pixel of a (global) graph: pixels whose upstream area exceeds
 *                        subtree_fraction * N / world form the trunk, the sub-trees hanging off it (and whole small
 *                        catchments) are bin-packed over the ranks, largest first; a trunk pixel joins the rank of its
 *                        largest tributary.  Deterministic: every rank computes the same map.  owner i32[N]
 *                        (compressed order, host or device); loads i64[world], n_trunk, n_roots may be NULL.
 *   lf_graph_cut_edges   the links (u -> d) of a graph whose ends have different owners, sorted by u: host arrays
 *                        edge_u / edge_d i32[cap] (compressed indices); *n_edges = total number found.
 *   lf_graph_restrict    the graph restricted to keep[N] != 0 (a rank's own pixels + the ghosts of the upstream ends
 *                        of its incoming cut edges).  Compressed indices of the result = rank of the pixel among the
 *                        kept ones; storage order and routing LEVELS stay the global ones, so the wavefront diagonals
 *                        of all ranks line up and sums over upstream pixels keep the reference's slot order: the cut
 *                        network reproduces the uncut one bit for bit.
 *   lf_xchg_*            one exchange region per rank (device memory: header + import_doubles float64 slots), made
 *                        visible to the other ranks through CUDA IPC (64-byte handles, exchanged by the host layer
 *                        over torch.distributed); the routing kernels store boundary discharges straight into the
 *                        consumer's region over NVLink and ghost pixels poll their slot (csrc/lf_xchg.cuh).
 *                        lf_xchg_set_peer_local wires two regions of ONE process to each other (single-GPU tests).
 *   lf_router_set_exchange / lf_model_set_exchange
 *                        xslot i32[N] (compressed order of the restricted graph): -1 plain pixel; k >= 0: upstream end
 *                        of outgoing cut edge k (export_peer[k] = consumer rank, export_offset[k] = offset in doubles of
 *                        (parity 0, that edge, section 0, step 0) inside the consumer's import area,
 *                        export_parity_stride[k] = doubles between the consumer's two parity copies); k <= -2: ghost of
 *                        incoming edge -2-k; INT32_MIN: ghost that has no link in this graph.  import_offset: offset in
 *                        doubles of this handle's import block in the own region; a block holds
 *                        [2 parities][n_import edges][sections][steps] with (sections, steps) = (1, cap_steps) for a
 *                        router, (3, 1) for the model's overland graph (which = 0) and (1 or 2, NoRoutSteps) for its
 *                        channel graph (which = 1).
 * ------------------------------------------------------------------------------------------- */
int lf_graph_partition(const lf_graph *g, int32_t world, double subtree_fraction, int32_t *owner, int64_t *loads,
                       int64_t *n_trunk, int64_t *n_roots);
int lf_graph_cut_edges(const lf_graph *g, const int32_t *owner, int64_t cap, int32_t *edge_u, int32_t *edge_d,
                       int64_t *n_edges);
int lf_graph_restrict(const lf_graph *g, const uint8_t *keep, lf_graph **out);
int lf_xchg_create(int32_t rank, int32_t world, int64_t import_doubles, lf_xchg **out);
int lf_xchg_ipc_handle(lf_xchg *x, void *handle64);
int lf_xchg_open_peer(lf_xchg *x, int32_t peer, const void *handle64);
int lf_xchg_set_peer_local(lf_xchg *x, int32_t peer, lf_xchg *other);
/* device address of a rank's region as seen from this process (peer = -1: the own region) */
int lf_xchg_peer_base(lf_xchg *x, int32_t peer, uint64_t *address);
/* Flow control around one run (a lf_router_run / a model step do this themselves): begin waits, on the device, until
 * every peer has finished the run before the previous one; end publishes the completion of this run. */
int lf_xchg_begin(lf_xchg *x, int32_t *parity);
int lf_xchg_end(lf_xchg *x);
/* aborted = 1 when a poll gave up (a peer stopped feeding its cut edges for ~10 s).  Synchronises. */
int lf_xchg_status(lf_xchg *x, int32_t *aborted, int64_t *epoch);
void lf_xchg_destroy(lf_xchg *x);
int lf_router_set_exchange(lf_router *r, lf_xchg *x, const int32_t *xslot, int32_t n_export, const int32_t *export_peer,
                           const int64_t *export_offset, const int64_t *export_parity_stride, int32_t n_import,
                           int64_t import_offset, int32_t cap_steps);
/* A model on two restricted graphs of the same pixel subset (LddToChan and LddKinematic); the model owns them. */
int lf_model_create_from_graphs(const lf_model_config *cfg, lf_graph *g_overland, lf_graph *g_channel, lf_model **out);
int lf_model_set_exchange(lf_model *m, lf_xchg *x, int32_t which, const int32_t *xslot, int32_t n_export,
                          const int32_t *export_peer, const int64_t *export_offset, const int64_t *export_parity_stride,
                          int32_t n_import, int64_t import_offset);

/* ---- The two Numba kernels of the reference as stand-alone operators (hydrological_modules/soilloop.py) ----
 * Same arguments, same in-place semantics; arrays are C-contiguous float64 (bool arrays: uint8), host or device
 * pointers.  (V,N) = (vegetation fraction, pixel), (L,N) = (land use, pixel), (N) = per pixel.  Host arrays are copied to
 * the device, updated there and copied back: the stand-alone operators exist for drop-in parity at the kernel level;
 * residency (and speed) is won one level up, in lf_model_soil. */

/* interception_water_balance(Interception, TaInterception, LeafDrainage, CumInterception, LAI, Rain, TaInterceptionMax,
 * drainageK), soilloop.py:27-70.  Writes Interception, TaInterception, LeafDrainage; updates CumInterception. */
int lf_interception_water_balance(double *Interception, double *TaInterception, double *LeafDrainage, double *CumInterception,
                                  const double *LAI, const double *Rain, const double *TaInterceptionMax, double drainageK,
                                  int64_t num_vegs, int64_t num_pixs);

/* suctionUnsaturatedSoilPF(index_landuse_all, pF0, pF1, pF2, W1a, W1b, W2, WRes*, WS*, PoreSpaceNotZero*, GenuInvAlpha*,
 * GenuInvM*, GenuInvN*, HeadMax), soilloop.py:402-424 (option simulatePF): pF = log10 of the capillary head [cm] of the
 * layers 1a, 1b, 2 -- the arrays below in that order --, -1 where the head is not positive.  Writes pF. */
typedef struct lf_soil_pf_args {
    int64_t num_vegs, num_pixs, num_landuses;
    const int64_t *index_landuse_all;   /* (V), host */
    double *pF[3];                      /* (V,N) out: pF0, pF1, pF2 */
    const double *W[3];                 /* (V,N): W1a, W1b, W2 */
    const double *WRes[3], *WS[3];      /* (L,N) */
    const uint8_t *PoreSpaceNotZero[3]; /* (L,N) */
    const double *GenuInvAlpha[3], *GenuInvM[3], *GenuInvN[3]; /* (L,N) */
    double HeadMax;
} lf_soil_pf_args;
int lf_suction_unsaturated_soil_pf(const lf_soil_pf_args *args);

/* soilColumnsWaterBalance(index_landuse_all, is_irrigated, is_paddy_irrig, paddy_inactive, DtDay, ...), soilloop.py:78-355:
 * the 73 arguments in the reference's order (paddy rice belongs to the EPIC crop module, out of scope: is_paddy_irrig
 * must be all false, paddy_inactive is ignored).  In/out: AvailableWaterForInfiltration, DSLR, ESAct, PrefFlow,
 * Infiltration, W1a, W1b, W1, W2, Theta*, Sat*, SeepTopToSubA/B, SeepSubToGW, UZOutflow, UZ, GwPercUZLZ. */
typedef struct lf_soil_columns_args {
    int64_t num_vegs, num_pixs, num_landuses;
    const int64_t *index_landuse_all; /* (V) */
    const uint8_t *is_irrigated;      /* (V) */
    const uint8_t *is_paddy_irrig;    /* (V), all 0 */
    double DtDay;
    double *AvailableWaterForInfiltration;                    /* (V,N) out */
    const double *Rain, *SnowMelt;                            /* (N) */
    const double *LeafDrainage, *Interception;                /* (V,N) */
    double *DSLR;                                             /* (V,N) in/out */
    double AvWaterThreshold;
    double *ESAct;                                            /* (V,N) out */
    const double *ESMax;                                      /* (V,N) */
    const uint8_t *isFrozenSoil;                              /* (N) */
    const double *b_Xinanjiang;                               /* (N) */
    const double *StoreMaxPervious;                           /* (L,N) */
    const double *PowerInfPot;                                /* (N) */
    double *PrefFlow;                                         /* (V,N) out */
    const double *PowerPrefFlow;                              /* (N) */
    double *Infiltration;                                     /* (V,N) out */
    double CourantCrit;
    const uint8_t *PoreSpaceNotZero1a, *PoreSpaceNotZero1b, *PoreSpaceNotZero2; /* (L,N) */
    const double *KSat1a, *KSat1b, *KSat2, *GenuInvM1a, *GenuInvM1b, *GenuInvM2, *GenuM1a, *GenuM1b, *GenuM2; /* (L,N) */
    double *W1a, *W1b, *W1, *W2;                              /* (V,N) in/out */
    double *Theta1a, *Theta1b, *Theta2, *Sat1a, *Sat1b, *Sat1, *Sat2; /* (V,N) out */
    double *SeepTopToSubA, *SeepTopToSubB, *SeepSubToGW;      /* (V,N) out */
    const double *WRes1a, *WRes1b, *WRes1, *WRes2, *WWP1a, *WWP1b, *WWP1, *WWP2, *WFC1a, *WFC1b, *WFC1, *WFC2; /* (L,N) */
    const double *SoilDepth1a, *SoilDepth1b, *SoilDepth2, *WS1a, *WS1b, *WS1, *WS2; /* (L,N) */
    const double *UpperZoneK;                                 /* (N) */
    double DrainedFraction;
    const double *GwPercStep;                                 /* (N) */
    double *UZOutflow, *UZ, *GwPercUZLZ;                      /* (V,N) out, in/out, out */
    int64_t *NoSubS_out;                                      /* optional (V,N): Darcy sub-steps taken */
} lf_soil_columns_args;
int lf_soil_columns_water_balance(const lf_soil_columns_args *args);

#ifdef __cplusplus
}
#endif
#endif /* LISFLOOD_B200_H */
