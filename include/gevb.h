/* =============================================================================
 * include/gevb.h -- C ABI of the B200-native gevolution hot path (libgevb.so)
 * =============================================================================
 * Drop-in boundary for the per-step particle-mesh path that gevolution 1.2's
 * main.cpp time loop (main.cpp:372-879) drives.  The reference has no FFI
 * layer: its "operator API" is the set of C++ free functions in gevolution.hpp
 * plus the LATfield2 methods they are called with.  Every entry point below
 * names the reference symbol (file:line) it replaces; include/gevolution_b200.hpp
 * re-exposes them under the reference's own C++ names and signatures.
 *
 * Conventions
 *   - plain C: opaque handles, pointers, sizes, doubles by value; no torch types.
 *   - every function returns 0 on success, non-zero on error (never exits);
 *     gevb_last_error() returns the message of the last failure on this thread.
 *   - all work is FP64 on one CUDA stream per context, asynchronous unless the
 *     call returns a scalar or copies to host memory.
 *   - one process drives one GPU (rank); ranks own z-slabs of the lattice:
 *     rank r owns planes [r*N/P, (r+1)*N/P) of every real field (+1 ghost plane
 *     below and above), the particles filed in those planes, and -- after a
 *     forward FFT when P > 1 -- the ky-slab [r*N/P, (r+1)*N/P) of Fourier space.
 *
 * Host-side array layouts (what upload/download exchange)
 *   real field      double[ncomp][nz_local][N][N]              [c][z][y][x]
 *   Fourier field   P == 1: double[ncomp][N][N][N/2+1][2]      [c][kz][ky][kx][re,im]
 *                   P  > 1: double[ncomp][nky_local][N][N/2+1][2]   [c][ky][kz][kx][re,im]
 *   particles       int64 id[n], double pos[n][3], double vel[n][3]
 *                   (vel = canonical momentum q/m, pos in box units [0,1))
 *   symmetric 3x3 tensor component order: (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
 * ============================================================================= */
#ifndef GEVB_H
#define GEVB_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gevb_ctx gevb_ctx;       /* LATfield2 `parallel` + `Lattice lat, latFT` (main.cpp:152,213-215) */
typedef struct gevb_field gevb_field;   /* LATfield2 Field<Real> / Field<Cplx>        (main.cpp:226-245)     */
typedef struct gevb_plan gevb_plan;     /* LATfield2 PlanFFT<Cplx>                    (main.cpp:238-246)     */
typedef struct gevb_pcls gevb_pcls;     /* Particles_gevolution<part_simple,...>      (main.cpp:217-219)     */

#define GEVB_REAL 0
#define GEVB_CPLX 1

#define GEVB_FFT_FORWARD 1              /* LATfield2 FFT_FORWARD  */
#define GEVB_FFT_BACKWARD (-1)          /* LATfield2 FFT_BACKWARD */

/* the particle callbacks main.cpp passes to updateVel / moveParticles; host
 * function pointers cannot run on the device, so the known ones are an enum  */
#define GEVB_UPDATE_Q 0                 /* update_q         gevolution.hpp:570 ; update_pos         :810 */
#define GEVB_UPDATE_Q_NEWTON 1          /* update_q_Newton  gevolution.hpp:709 ; update_pos_Newton  :900 */
#define GEVB_DISPLACE_PCLS_IC_BASIC 2   /* displace_pcls_ic_basic  ic_basic.hpp:60  (moveParticles only)  */
#define GEVB_INITIALIZE_Q_IC_BASIC 3    /* initialize_q_ic_basic   ic_basic.hpp:118 (updateVel only)      */

/* ---- errors / versioning -------------------------------------------------- */
const char * gevb_last_error(void);
const char * gevb_version(void);

/* kernel-variant knobs for ablation runs; results do not depend on them (first value = default):
 *   "geodesic_variant"  block shape / prefetch of the kick-drift kernel, see geodesic.cu
 *   "fft_exchange"      1 = transposes pushed over peer memory, 0 = NCCL all-to-all + local transpose
 *   "fft_overlap"       pushes overlap the local transforms: 2 = piece by piece forward and component by component backward,
 *                       3 = piece by piece both ways, 1 = component by component, 0 = not at all
 *   "fft_decomposed"    1 = single-rank transforms as 2-D per plane + 1-D along z, 0 = cuFFT 3-D plans
 *   "fft_l2_planes"     0 = off, n = 2-D passes of a component in chunks of n planes
 *   "deposit_variant"   4 = site tile flushed by bulk reductions, 0 = by one RED per site, 1 = per-cell accumulators in shared
 *                       memory, 2 = thread per cell with register accumulators (deposit.cu)
 *   "rebin_variant"     2 = re-bin move through a shared-memory window, 0 = direct scattered stores; +1 = slots from ranks the
 *                       drift kernel recorded instead of counting the histogram down (particles.cu)
 *   "peer_comm"         1 = halo / fold / migration over peer memory with flag barriers, 0 = NCCL point-to-point
 *   "geodesic_tma"      1 = field tiles of the kick/drift kernel by TMA tensor loads (bricks away from the lattice edge), 0 = LDGSTS
 *   "tma_l2_promotion"  L2 promotion of the tensor maps: 0 = none, 1 / 2 / 3 = 64 / 128 / 256 bytes
 *   "fft_xpass"         0 = cuFFT only, 1 = forward transforms at N = 512 on one rank through the own x-pass (xpass.cu, prepareFTsource
 *                       fused into its load) + one strided 2-D cuFFT pass -- correct, measured slower */
int gevb_tuning(const char * knob, int value);

/* ---- context: lattice geometry + device + communicator --------------------
 * replaces parallel.initialize(n,m) (main.cpp:152), Lattice lat(3,box,halo)
 * (main.cpp:213) and latFT.initializeRealFFT (main.cpp:215).
 * nccl_id: NULL when nranks == 1, else the 128-byte ncclUniqueId produced by
 * gevb_nccl_unique_id() on rank 0 and distributed by the host program.       */
int gevb_nccl_unique_id(void * out128);
int gevb_ctx_create(gevb_ctx ** out, int ngrid, int device, int rank, int nranks, const void * nccl_id);
/* the slab decomposition gevb_ctx_create uses, by itself (no device needed): z-planes of real space and ky-rows of
 * Fourier space owned by `rank` */
int gevb_slab_geometry(int ngrid, int rank, int nranks, int * z0, int * nz_local, int * ky0, int * nky_local);
int gevb_ctx_destroy(gevb_ctx * ctx);
int gevb_ctx_sync(gevb_ctx * ctx);                         /* cudaStreamSynchronize                */
int gevb_ctx_geometry(gevb_ctx * ctx, int * ngrid, int * z0, int * nz_local, int * ky0, int * nky_local);
void * gevb_ctx_stream(gevb_ctx * ctx);                    /* cudaStream_t, for event timing       */
/* kernels of this library launched on ctx since creation (bench.py gpu_launches) */
int64_t gevb_ctx_launch_count(gevb_ctx * ctx);
/* optional per-call device timing (the reference's -DBENCHMARK timers, main.cpp:71-88, per entry point):
 * enable, run, then read milliseconds and call counts per class since the last read            */
int gevb_ctx_timing(gevb_ctx * ctx, int enable);
int gevb_ctx_timing_read(gevb_ctx * ctx, double * ms, int64_t * counts);   /* arrays of gevb_timing_num_classes() */
int gevb_timing_num_classes(void);
const char * gevb_timing_class_name(int cls);
/* parallel.sum / parallel.max (main.cpp:462,816): all-reduce n doubles in place (host values) */
int gevb_parallel_sum(gevb_ctx * ctx, double * v, int n);
int gevb_parallel_max(gevb_ctx * ctx, double * v, int n);

/* ---- fields ------------------------------------------------------------------
 * Field::initialize(lat, ncomp) / (lat,3,3,symmetric) + alloc (main.cpp:234-245) */
int gevb_field_create(gevb_ctx * ctx, gevb_field ** out, int kind, int ncomp, int symmetric);
int gevb_field_destroy(gevb_field * f);
int gevb_field_upload(gevb_field * f, const double * host);      /* bulk only; ghost planes untouched */
int gevb_field_download(gevb_field * f, double * host);
int gevb_field_components(gevb_field * f);
void * gevb_field_device_ptr(gevb_field * f);
/* projection_init(Field*) (main.cpp:378,426,438): zero all components incl. ghost planes */
int gevb_projection_init(gevb_field * f);
/* Field::updateHalo() (main.cpp:518,568,598): periodic ghost fill (z planes; x,y wrap is index math) */
int gevb_field_updateHalo(gevb_field * f);
/* scalarProjectionCIC_comm / vectorProjectionCICNGP_comm / symtensorProjectionCICNGP_comm
 * (gevolution.hpp:1024,1149,1300; main.cpp:411,435,450): fold the upper ghost plane into
 * the periodically next rank's first bulk plane                                            */
int gevb_projection_comm(gevb_field * f);
/* sum over the local bulk of one component, then parallel.sum (main.cpp:459-463) */
int gevb_field_sum(gevb_field * f, int comp, double * out);
/* result(x) += value on the bulk of comp (the bg_ncdm add, main.cpp:394-397) */
int gevb_field_add_constant(gevb_field * f, int comp, double value);
/* every component of the local bulk times factor: the site loops that rescale the stored vector potential around an
 * output or a restart (output.hpp:212-218, hibernation.hpp:533-538, ic_read.hpp:312-317)                             */
int gevb_field_scale(gevb_field * f, double factor);
gevb_ctx * gevb_field_ctx(gevb_field * f);
/* Field::saveHDF5 / loadHDF5 (output.hpp:98-300, hibernation.hpp:588-600, ic_read.hpp:310,326): HDF5 is not in this
 * image, so the dataset is written as a flat binary file -- 32-byte header {"GEVBFLD1", int32 ngrid, ncomp, 16 bytes 0},
 * then float64 [comp][z][y][x] of the whole lattice.  Collective: every rank writes / reads the planes of its slab.   */
int gevb_field_save_raw(gevb_field * f, const char * filename);
int gevb_field_load_raw(gevb_field * f, const char * filename);

/* ---- FFT ---------------------------------------------------------------------
 * PlanFFT<Cplx>(&real,&cplx) + execute(dir) (main.cpp:238-246,477,488,544,563,575,593):
 * per-component 3-D r2c / c2r, unnormalised both ways; the Fourier input of a
 * backward transform is preserved (BiFT is persistent state, main.cpp:586-593). */
int gevb_plan_create(gevb_plan ** out, gevb_field * real_field, gevb_field * cplx_field);
int gevb_plan_destroy(gevb_plan * plan);
int gevb_plan_execute(gevb_plan * plan, int direction);
/* preserve = 0: a backward execute may clobber the plan's Fourier field (saves one copy of it per execute).
 * The main loop does this for plan_phi / plan_chi, whose scalarFT is scratch that the next forward
 * transform overwrites (main.cpp:477-488,558-563).  Default 1 (LATfield2 semantics).               */
int gevb_plan_set_preserve_input(gevb_plan * plan, int preserve);

/* ---- particles ---------------------------------------------------------------
 * Particles::initialize + addParticle_global (ic_basic.hpp:1990,1429): particles
 * whose cell lies in this rank's slab are kept, the rest ignored.  Storage is
 * cell-sorted FP64 structure-of-arrays (bricks of cells, then cells inside the
 * brick); cell = floor(pos/dx) per axis.                                      */
int gevb_pcls_create(gevb_ctx * ctx, gevb_pcls ** out, double mass);
int gevb_pcls_destroy(gevb_pcls * p);
/* Particles::initialize on a container that already exists: drops its particles, keeps the device arrays */
int gevb_pcls_reset(gevb_pcls * p, double mass);
int gevb_pcls_add(gevb_pcls * p, int64_t n, const int64_t * id, const double * pos, const double * vel);
int gevb_pcls_count(gevb_pcls * p, int64_t * n_local);
int gevb_pcls_download(gevb_pcls * p, int64_t * id, double * pos, double * vel);   /* storage (brick-major cell) order */
/* bit-exact contract: particles per cell of the local slab, uint32[nz_local][N][N] */
int gevb_pcls_cell_counts(gevb_pcls * p, uint32_t * counts);
double gevb_pcls_mass(gevb_pcls * p);
/* storage order of the particle arrays (all indices z-major): super-bricks of super3[] bricks, then the brick
 * inside the super-brick, then the cell inside the brick of brick3[] cells:
 * key = ((super * bricks_per_super + brick) * cells_per_brick) + cell                                       */
void gevb_brick_dims(int * brick3, int * super3);

/* ---- particle -> mesh projections (gevolution.hpp:927,1046,1173; main.cpp:385,402,427,439)
 * phi may be NULL (no geometric correction, gevolution.hpp:949,965).  Target
 * fields accumulate (several species project into one field); the target's
 * ghost planes must be valid targets (projection_init zeroes them).            */
int gevb_projection_T00_project(gevb_pcls * p, gevb_field * T00, double a, gevb_field * phi, double coeff);
int gevb_projection_T0i_project(gevb_pcls * p, gevb_field * T0i, gevb_field * phi, double coeff);
int gevb_projection_Tij_project(gevb_pcls * p, gevb_field * Tij, double a, gevb_field * phi, double coeff);
int gevb_scalarProjectionCIC_project(gevb_pcls * p, gevb_field * rho);
/* one pass over the particles for T00 and Tij together (same results as the two calls) */
int gevb_projection_T00_Tij_project(gevb_pcls * p, gevb_field * T00, gevb_field * Tij, double a, gevb_field * phi, double coeff);

/* ---- real-space source preparation (gevolution.hpp:57,170; main.cpp:472,539);
 * result may alias source / Sij may alias Tij                                    */
int gevb_prepareFTsource_scalar(gevb_field * phi, gevb_field * chi, gevb_field * source, double bgmodel, gevb_field * result, double coeff, double coeff2, double coeff3);
/* same, and also returns sum(source) over the lattice before it is modified -- the T00hom sum of main.cpp:459-462
 * (local sum + parallel.sum) without its own pass over the field                     */
int gevb_prepareFTsource_scalar_sum(gevb_field * phi, gevb_field * chi, gevb_field * source, double bgmodel, gevb_field * result, double coeff, double coeff2, double coeff3, double * sum_source);
int gevb_prepareFTsource_tensor(gevb_field * phi, gevb_field * Tij, gevb_field * Sij, double coeff);
/* fused forms of main.cpp:472+477 and :539+544: prepareFTsource(...) in place on the plan's real field followed by
 * plan.execute(FFT_FORWARD).  Where the own x-pass is switched on (knob fft_xpass; one rank, N = 512) the preparation rides on the load of
 * the transform's first pass and the prepared values are never stored: on return the plan's Fourier field holds the transform of
 * the prepared source, the real field still holds the incoming one.  Elsewhere: the two calls one after the other.
 * sum_source may be NULL.                                                          */
int gevb_prepareFTsource_scalar_fft(gevb_field * phi, gevb_field * chi, gevb_plan * plan_source, double bgmodel, double coeff, double coeff2, double coeff3, double * sum_source);
int gevb_prepareFTsource_tensor_fft(gevb_field * phi, gevb_plan * plan_Sij, double coeff);

/* ---- Fourier-space kernels (gevolution.hpp:211,284,350,411,501); outputs may alias inputs */
int gevb_solveModifiedPoissonFT(gevb_field * sourceFT, gevb_field * potFT, double coeff, double modif);
int gevb_projectFTscalar(gevb_field * SijFT, gevb_field * chiFT, int add);
int gevb_evolveFTvector(gevb_field * SijFT, gevb_field * BiFT, double a2dtau);
/* projectFTscalar (main.cpp:558) and evolveFTvector (:586) on one read of SijFT: same results as the two calls */
int gevb_projectFTscalar_evolveFTvector(gevb_field * SijFT, gevb_field * chiFT, gevb_field * BiFT, double a2dtau);
int gevb_projectFTvector(gevb_field * SiFT, gevb_field * BiFT, double coeff, double modif);
int gevb_projectFTtensor(gevb_field * SijFT, gevb_field * hijFT);

/* ---- geodesic updates ----------------------------------------------------------
 * Particles::updateVel(fn, dtau, fields, nfields, params) (main.cpp:775): returns
 * sqrt(max v^2) over the LOCAL particles in *maxvel (caller reduces, main.cpp:816).
 * Particles::moveParticles(fn, dtau, fields, nfields, params) (main.cpp:798):
 * drift, periodic wrap, re-bin (cell-sorted order restored), slab migration.
 * fields = {phi, chi, Bi} with valid ghost planes; params = {a, a^2 N}.         */
int gevb_updateVel(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params, double * maxvel);
int gevb_moveParticles(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params);
/* moveParticles with the callback's reduction output (LATfield2's output / reduce_type / noutput arguments, used
 * by the IC generator: moveParticles(displace_pcls_ic_basic, 1., fields, n, NULL, &max_displacement, &MAX, 1),
 * ic_basic.hpp:1995): *output_max = largest displacement of the LOCAL particles (caller reduces)              */
int gevb_moveParticles_max(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params, double * output_max);
/* fused kick (main.cpp:775) + drift (main.cpp:798) in one pass over the particles: positions
 * do not change between the two reference calls, only params (a advances by rungekutta4bg,
 * main.cpp:792), so the result equals updateVel followed by moveParticles.      */
int gevb_kick_drift(gevb_pcls * p, int fn, double dtau_kick, int nfields_kick, const double * params_kick,
                    double dtau_drift, int nfields_drift, const double * params_drift,
                    gevb_field * const * fields, double * maxvel);

/* ---- analysis (tools.hpp:53,237): binned power spectrum incl. the final
 * normalisation of tools.hpp:186-193 (all ranks receive the result)             */
int gevb_extractPowerSpectrum(gevb_field * fldFT, double * kbin, double * power, double * kscatter, double * pscatter, int * occupation, int numbins, int deconvolve, int ktype);

/* tools.hpp:268-346: the reference's power-spectrum text file (rank 0 calls it), including the interpolation to the
 * exact output redshift when the file of the previous cycle exists (EXACT_OUTPUT_REDSHIFTS); z_target < 0 disables it */
int gevb_writePowerSpectrum(const double * kbin, const double * power, const double * kscatter, const double * pscatter, const int * occupation, int numbins,
                            double rescalek, double rescalep, const char * filename, const char * description, double a, double z_target);

/* ---- snapshots (Particles_gevolution.hpp:30-251) ---------------------------------
 * saveGadget2: Gadget-2 binary file of one species, float32 positions [kpc/h] / velocities [km/s / sqrt(a)], int64
 * IDs, tracers ID % tracer_factor == 0, with the half-step corrections of EXACT_OUTPUT_REDSHIFTS (dtau_pos, dtau_vel,
 * phi may be NULL).  header256 is the caller's 256-byte gadget2_header (metadata.hpp:152-171); its npart[1] is set
 * to the number written.  Collective: every rank writes its share of the one file.
 * gevb_pcls_gadget2_arrays is the device half (selection, corrections, units) returning host arrays.            */
int gevb_pcls_saveGadget2(gevb_pcls * p, const char * filename, void * header256, int tracer_factor, double dtau_pos, double dtau_vel, gevb_field * phi);
int gevb_pcls_gadget2_arrays(gevb_pcls * p, double a, double boxsize, int tracer_factor, double dtau_pos, double dtau_vel, gevb_field * phi,
                             float * pos, float * vel, int64_t * ids, int64_t * n_out);
gevb_ctx * gevb_pcls_ctx(gevb_pcls * p);
int gevb_ctx_ranks(gevb_ctx * ctx, int * rank, int * nranks);

/* ---- time loop (main.cpp:372-879, outputs stripped), host side in C++ ---------
 * dsettings: boxsize, Cf, steplimit, z_in, z_relax
 * cosmo    : Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, w0_fld, wa_fld, Omega_g, Omega_ur, Omega_rad, h */
typedef struct gevb_sim gevb_sim;
/* host-side background model by itself, no device needed (background.hpp:103,137,167,200):
 * out4 = Hconf(a), bg_ncdm(a), particleHorizon(a), a after rungekutta4bg(a, dtau) */
int gevb_background_eval(const double * cosmo, int num_ncdm, const double * m_ncdm, const double * T_ncdm, const double * Omega_ncdm,
                         double a, double fourpiG, double dtau, double * out4);
int gevb_sim_create(gevb_sim ** out, gevb_ctx * ctx, int gr_flag, int vector_flag, const double * dsettings, const double * cosmo);
int gevb_sim_destroy(gevb_sim * sim);
/* non-cold dark matter species (at most 4 = MAX_PCL_SPECIES-2, metadata.hpp:48): masses [eV], temperatures
 * [T_cmb], density parameters (metadata.hpp:284-288), the redshifts at which their T00 / T0i deposits switch
 * on ("switch delta_ncdm", "switch B ncdm", parser.hpp:1755-1793), "switch linear chi" (Tij) and the move limit
 * that sets the sub-cycling of their updates (main.cpp:696-701).  Their particles are species 2..5.          */
int gevb_sim_set_ncdm(gevb_sim * sim, int num_ncdm, const double * m_ncdm, const double * T_ncdm, const double * Omega_ncdm,
                      const double * z_switch_deltancdm, const double * z_switch_Bncdm, double z_switch_linearchi, double movelimit);
int gevb_sim_set_ncdm_maxvel(gevb_sim * sim, const double * maxvel4);          /* max |v| per ncdm species as the IC generator returns it */
int gevb_sim_get_ncdm_state(gevb_sim * sim, double * maxvel4, int * numsteps4); /* after a cycle: max |v| and step subdivisions used  */
/* species: 0 cdm, 1 baryons, 2..5 ncdm */
int gevb_sim_set_particles(gevb_sim * sim, int species, int64_t n, const int64_t * id, const double * pos, const double * vel, double mass);
int gevb_sim_set_field(gevb_sim * sim, int which, const double * host);   /* 0 phi,1 chi,2 Bi,3 source,4 Sij,10 scalarFT,11 BiFT,12 SijFT */
int gevb_sim_get_field(gevb_sim * sim, int which, double * host);
gevb_field * gevb_sim_field(gevb_sim * sim, int which);
gevb_pcls * gevb_sim_pcls(gevb_sim * sim, int species);
int gevb_sim_get_state(gevb_sim * sim, double * out9);     /* a,tau,dtau,dtau_old,cycle,maxvel0,maxvel1,T00hom,fourpiG */
int gevb_sim_set_state(gevb_sim * sim, const double * in7);
int gevb_sim_set_fused(gevb_sim * sim, int fused);          /* 1 (default): fused deposit + fused kick/drift, scalarFT is scratch after a cycle; 0: one call per reference call */
/* writeSpectra for phi, chi, hij, B (output.hpp:1945-1981,2151-2155): mask = MASK_PHI 1 | MASK_CHI 2 | MASK_B 8 |
 * MASK_HIJ 128 (metadata.hpp:56-63); files <prefix><pkcount %03d>_<phi|chi|hij|B>.dat                         */
int gevb_sim_write_spectra(gevb_sim * sim, const char * prefix, int pkcount, int numbins, int mask, double z_target);
/* writeSnapshots' Gadget-2 branch for one species (output.hpp:95-131 header, then saveGadget2) */
int gevb_sim_save_gadget2(gevb_sim * sim, int species, const char * filename, int tracer_factor, double dtau_pos, double dtau_vel);
/* hibernation / restart with the reference's state set and arithmetic (hibernation.hpp:512-611, ic_read.hpp:290-330):
 * the particles of every species (<filebase>.<rank>.gevb, with the loop scalars a, tau, dtau, dtau_old, cycle, maxvel),
 * phi and chi (<filebase>_phi.bin, _chi.bin) and the vector potential in REAL space divided by a^2 N (<filebase>_B.bin,
 * hibernation.hpp:533-538).  Restore multiplies B by a^2 / N^2 and rebuilds BiFT by a forward transform
 * (ic_read.hpp:305-319).  HDF5 is not in this image: the field files are flat binary (gevb_field_save_raw).  Restore
 * needs a sim created with the same lattice, decomposition and flags.  Both are collective.                     */
/* writeSnapshots' field dumps (output.hpp:98-300) as raw binary files <prefix>_<T00|B|phi|chi|hij>.bin (see
 * gevb_field_save_raw): mask = MASK_PHI 1 | MASK_CHI 2 | MASK_B 8 | MASK_T00 16 | MASK_HIJ 128.  B is written divided by
 * a^2 N and restored afterwards by a backward transform of BiFT, exactly the reference's sequence (output.hpp:212-236);
 * hij is the TT projection of SijFT transformed back (output.hpp:259-265), T00 a fresh projection (:155-182).          */
int gevb_sim_write_field_snapshot(gevb_sim * sim, const char * prefix, int mask);
int gevb_sim_hibernate(gevb_sim * sim, const char * filebase);
int gevb_sim_restore(gevb_sim * sim, const char * filebase);
/* the main loop with its power-spectrum and Gadget-2 snapshot outputs at the requested redshifts (main.cpp:372-879:
 * output scheduling :617-679 with the EXACT_OUTPUT_REDSHIFTS logic, termination :685-693); files
 * <pk_prefix><count %03d>_<phi|chi|hij|B>.dat and <snap_prefix><count %03d>_cdm (_b, _ncdm<i>).  Redshift lists in
 * descending order as the parser leaves them.  counts3 (may be NULL) = cycles run, spectra sets, snapshots written. */
int gevb_sim_run(gevb_sim * sim, const double * z_pk, int num_pk, int pk_mask, int numbins, const char * pk_prefix,
                 const double * z_snapshot, int num_snapshot, int tracer_factor, const char * snap_prefix, int max_cycles, int * counts3);
int gevb_sim_step(gevb_sim * sim);                          /* one cycle; asynchronous except the maxvel / T00hom reads */

/* ---- settings.ini and the basic IC generator: a run starts from the reference's own settings file ---------------
 * gevb_settings_read restates the subset of the reference's parser (parser.hpp:40-105 readline, :122 loadParameterFile,
 * :759-1800 parseMetadata) that the hot path, its outputs and "IC generator = basic" need; keys, defaults and derived
 * cosmological parameters are the reference's.  `overrides` (may be NULL) holds further "key = value" lines that replace
 * the file's lines of the same key.  Not supported (reported as errors): mPk file, IC generator other than basic,
 * ncdm particle species from the generator, CLASS, lightcones.                                                       */
#define GEVB_MAX_OUTPUTS 32
#define GEVB_PATH_MAX 512
typedef struct gevb_settings
{
	int ngrid, gr_flag, vector_flag;        /* Ngrid; gravity theory GR = 1 / Newton = 0; vector method parabolic = 0 / elliptic = 1 */
	int baryon_flag;                        /* baryon treatment: ignore 0, sample 1, blend 2, hybrid 3 (parser.hpp:928-960)            */
	int seed, ksphere, correct_displacement;/* seed; k-domain sphere; correct displacement                                             */
	int tiling[2];                          /* tiling factor of the cdm (and baryon) template                                          */
	int tracer_factor[2];
	int numbins, pk_mask, snapshot_mask;    /* Pk bins; Pk outputs / snapshot outputs as MASK_* bits (metadata.hpp:56-70)              */
	int num_pk, num_snapshot;
	double boxsize, Cf, steplimit, movelimit, z_in, z_relax;
	double A_s, n_s, k_pivot;
	double cosmo[11];                       /* Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, w0_fld, wa_fld, Omega_g, Omega_ur, Omega_rad, h */
	double z_pk[GEVB_MAX_OUTPUTS], z_snapshot[GEVB_MAX_OUTPUTS];   /* descending, as the parser leaves them                           */
	char template_file[2][GEVB_PATH_MAX], tk_file[GEVB_PATH_MAX];
	char output_path[GEVB_PATH_MAX], basename_generic[128], basename_pk[128], basename_snapshot[128];
} gevb_settings;
int gevb_settings_read(const char * filename, const char * overrides, gevb_settings * out);
/* main.cpp:184-340 for IC generator = basic: a simulation with the file's lattice, flags and cosmology (ctx must have been
 * created with settings->ngrid), particles and metric fields from generateIC_basic (ic_basic.hpp:1626-2259; Threefry
 * realisation prng_engine.hpp, transfer-function splines, template tiling, displacement / velocity callbacks, phi, chi, B). */
int gevb_sim_create_from_settings(gevb_sim ** out, gevb_ctx * ctx, const gevb_settings * settings);
/* the main loop with the file's outputs (gevb_sim_run with the settings' redshift lists, masks and file names under
 * output_path)                                                                                                          */
int gevb_sim_run_settings(gevb_sim * sim, const gevb_settings * settings, int max_cycles, int * counts3);
/* host-side pieces of the generator by themselves (no device): the CIC convolution kernel on its 3 x 3 x 3 support around the
 * origin (generateCICKernel, ic_basic.hpp:737-1052; out27[(dz+1)*9 + (dy+1)*3 + (dx+1)], numpcl = 0 gives the standard kernel)
 * and the Gaussian realisation of one Fourier field (generateDisplacementField, ic_basic.hpp:1090-1379) on the host layout
 * double[kz][ky][kx][2] of a whole lattice; potFT holds the kernel's transform on entry.                                */
int gevb_ic_cic_kernel(int ngrid, int64_t numpcl, const float * pcldata, int numtile, double * out27);
int gevb_ic_displacement_field(int ngrid, double * potFT, double coeff, int nspline, const double * spline_x, const double * spline_y,
                               unsigned int seed, int ksphere, int deconvolve_f);
/* loadHomogeneousTemplate (ic_basic.hpp:191-330): positions of a Gadget-2 template in box units; *numpart particles, pcldata
 * (3 floats each) is allocated with malloc and owned by the caller                                                      */
int gevb_ic_load_template(const char * filename, int64_t * numpart, float ** pcldata);
/* sets individual sites of one component of a real field (global coordinates; sites outside this rank's slab are skipped) */
int gevb_field_set_sites(gevb_field * f, int comp, int n, const int * xyz, const double * values);

#ifdef __cplusplus
}
#endif
#endif
