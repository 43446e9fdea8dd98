// =============================================================================
// include/gevolution_b200.hpp -- C++ drop-in surface over the C ABI (gevb.h)
// =============================================================================
// Re-exposes the hot path under the reference's own names and argument order so
// that a main.cpp-shaped time loop (reference main.cpp:372-879) compiles against
// device-resident handles instead of LATfield2 objects:
//
//   reference (LATfield2 / gevolution.hpp)                 here
//   ------------------------------------------------------------------------
//   parallel.initialize(n,m); parallel.sum/max  main.cpp:152,462,816   Parallel
//   Lattice lat(3,box,halo); latFT              main.cpp:213-215       Lattice
//   Field<Real>, Field<Cplx>                    main.cpp:226-245       Field<Real>, Field<Cplx>
//   PlanFFT<Cplx> plan(&real,&cplx); execute    main.cpp:238-246,477   PlanFFT<Cplx>
//   Particles_gevolution<part_simple,...>       main.cpp:217-219       Particles_gevolution
//   projection_init / *_project / *_comm        main.cpp:378-450       same names
//   prepareFTsource, solveModifiedPoissonFT,
//   projectFTscalar/vector/tensor, evolveFTvector  gevolution.hpp      same names
//   pcls.updateVel(update_q, ...)               main.cpp:775           same (callback identity -> enum)
//   pcls.moveParticles(update_pos, ...)         main.cpp:798           same
//
// Errors: the reference prints and exit(-1)s / parallel.abortForce()s; the C ABI
// returns a status.  This wrapper converts a non-zero status into the
// reference's behaviour (message on stderr, abort).
// =============================================================================
#ifndef GEVOLUTION_B200_HPP
#define GEVOLUTION_B200_HPP

#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "gevb.h"

namespace gevb200 {

typedef double Real;
struct Cplx { Real re, im; };

#define FFT_FORWARD GEVB_FFT_FORWARD
#define FFT_BACKWARD GEVB_FFT_BACKWARD

// define GEVB_THROW_ON_ERROR to get a C++ exception instead of the reference's abort
struct gevb_error { const char * what; };
inline void check(int status, const char * what)
{
	if (status != 0)
	{
#ifdef GEVB_THROW_ON_ERROR
		gevb_error e = {what};
		throw e;
#else
		std::fprintf(stderr, " error in %s: %s\n", what, gevb_last_error());
		std::abort();                       // parallel.abortForce()
#endif
	}
}

enum Symmetry { unsymmetric = 0, symmetric = 1 };

// Lattice + parallel: geometry, device and communicator of one rank
class Lattice
{
	gevb_ctx * ctx_;
	bool owner_;
public:
	Lattice() : ctx_(NULL), owner_(false) {}
	explicit Lattice(gevb_ctx * ctx) : ctx_(ctx), owner_(false) {}
	~Lattice() { if (owner_ && ctx_) gevb_ctx_destroy(ctx_); }
	void initialize(int ngrid, int device = 0, int rank = 0, int nranks = 1, const void * nccl_id = NULL)
	{
		check(gevb_ctx_create(&ctx_, ngrid, device, rank, nranks, nccl_id), "Lattice::initialize");
		owner_ = true;
	}
	gevb_ctx * ctx() const { return ctx_; }
	int size(int) const { int n; gevb_ctx_geometry(ctx_, &n, NULL, NULL, NULL, NULL); return n; }
	int halo() const { return 1; }
	// parallel.sum / parallel.max (main.cpp:462,816)
	void sum(double * v, int n) { check(gevb_parallel_sum(ctx_, v, n), "parallel.sum"); }
	void max(double * v, int n) { check(gevb_parallel_max(ctx_, v, n), "parallel.max"); }
};

template <class T> struct FieldKind;
template <> struct FieldKind<Real> { enum { kind = GEVB_REAL }; };
template <> struct FieldKind<Cplx> { enum { kind = GEVB_CPLX }; };

template <class T>
class Field
{
	gevb_field * f_;
	Lattice * lat_;
	Field(const Field &);
	Field & operator=(const Field &);
public:
	Field() : f_(NULL), lat_(NULL) {}
	~Field() { if (f_) gevb_field_destroy(f_); }
	void initialize(Lattice & lat, int components = 1)
	{
		lat_ = &lat;
		check(gevb_field_create(lat.ctx(), &f_, FieldKind<T>::kind, components, 0), "Field::initialize");
	}
	void initialize(Lattice & lat, int rows, int cols, Symmetry sym)
	{
		lat_ = &lat;
		check(gevb_field_create(lat.ctx(), &f_, FieldKind<T>::kind, sym == symmetric ? rows * (rows + 1) / 2 : rows * cols, sym == symmetric), "Field::initialize");
	}
	void alloc() {}
	gevb_field * handle() const { return f_; }
	Lattice & lattice() const { return *lat_; }
	int components() const { return gevb_field_components(f_); }
	void updateHalo() { check(gevb_field_updateHalo(f_), "Field::updateHalo"); }
	void upload(const double * host) { check(gevb_field_upload(f_, host), "Field::upload"); }
	void download(double * host) { check(gevb_field_download(f_, host), "Field::download"); }
};

template <class T>
class PlanFFT
{
	gevb_plan * p_;
public:
	PlanFFT() : p_(NULL) {}
	PlanFFT(Field<Real> * r, Field<T> * k) : p_(NULL) { initialize(r, k); }
	~PlanFFT() { if (p_) gevb_plan_destroy(p_); }
	void initialize(Field<Real> * r, Field<T> * k) { check(gevb_plan_create(&p_, r->handle(), k->handle()), "PlanFFT"); }
	void execute(int direction) { check(gevb_plan_execute(p_, direction), "PlanFFT::execute"); }
	// extension: a backward execute may clobber the Fourier field (it is scratch for the caller)
	gevb_plan * handle() const { return p_; }
	void preserveInput(bool keep) { check(gevb_plan_set_preserve_input(p_, keep ? 1 : 0), "PlanFFT::preserveInput"); }
};

struct part_simple { long ID; Real pos[3]; Real vel[3]; };
struct part_simple_info { double mass; int relativistic; char type_name[64]; };
struct Site;   // never dereferenced: present only so the callback signatures match

// the reference's particle callbacks; on the device they are selected by identity, never called
typedef Real (*updateVel_fn)(double, double, part_simple *, double *, part_simple_info, Field<Real> **, Site *, int, double *, double *, int);
typedef void (*moveParticles_fn)(double, double, part_simple *, double *, part_simple_info, Field<Real> **, Site *, int, double *, double *, int);
inline Real update_q(double, double, part_simple *, double *, part_simple_info, Field<Real> **, Site *, int, double *, double *, int) { return 0.; }          // gevolution.hpp:570
inline Real update_q_Newton(double, double, part_simple *, double *, part_simple_info, Field<Real> **, Site *, int, double *, double *, int) { return 0.; }   // gevolution.hpp:709
inline void update_pos(double, double, part_simple *, double *, part_simple_info, Field<Real> **, Site *, int, double *, double *, int) {}                   // gevolution.hpp:810
inline void update_pos_Newton(double, double, part_simple *, double *, part_simple_info, Field<Real> **, Site *, int, double *, double *, int) {}            // gevolution.hpp:900
inline void displace_pcls_ic_basic(double, double, part_simple *, double *, part_simple_info, Field<Real> **, Site *, int, double *, double *, int) {}       // ic_basic.hpp:60
inline Real initialize_q_ic_basic(double, double, part_simple *, double *, part_simple_info, Field<Real> **, Site *, int, double *, double *, int) { return 0.; }   // ic_basic.hpp:118
#define MAX 1                              // LATfield2 reduction type of the callback outputs (only MAX is used, ic_basic.hpp:1994)

class Particles_gevolution
{
	gevb_pcls * p_;
	Lattice * lat_;
	static void handles(Field<Real> ** fields, int nfields, gevb_field ** out) { for (int i = 0; i < nfields && i < 3; i++) out[i] = fields[i]->handle(); }
public:
	Particles_gevolution() : p_(NULL), lat_(NULL) {}
	~Particles_gevolution() { if (p_) gevb_pcls_destroy(p_); }
	void initialize(part_simple_info info, Lattice * lat)
	{
		if (p_ && lat_ == lat) { check(gevb_pcls_reset(p_, info.mass), "Particles::initialize"); return; }   // same lattice: keep the device arrays
		if (p_) { gevb_pcls_destroy(p_); p_ = NULL; }
		lat_ = lat;
		check(gevb_pcls_create(lat->ctx(), &p_, info.mass), "Particles::initialize");
	}
	bool initialized() const { return p_ != NULL; }
	gevb_pcls * handle() const { return p_; }
	// bulk form of addParticle_global (ic_basic.hpp:1429)
	void addParticles_global(int64_t n, const int64_t * id, const double * pos, const double * vel) { check(gevb_pcls_add(p_, n, id, pos, vel), "Particles::addParticle_global"); }
	int64_t numParticlesLocal() const { int64_t n = 0; gevb_pcls_count(p_, &n); return n; }
	Real updateVel(updateVel_fn fn, double dtau, Field<Real> ** fields, int nfields, double * params = NULL)
	{
		int kind = -1;
		if (fn == &update_q) kind = GEVB_UPDATE_Q; else if (fn == &update_q_Newton) kind = GEVB_UPDATE_Q_NEWTON; else if (fn == &initialize_q_ic_basic) kind = GEVB_INITIALIZE_Q_IC_BASIC;
		gevb_field * h[3] = {NULL, NULL, NULL};
		handles(fields, nfields, h);
		double maxvel = 0.;
		check(gevb_updateVel(p_, kind, dtau, h, nfields, params, &maxvel), "Particles::updateVel");
		return maxvel;
	}
	void moveParticles(moveParticles_fn fn, double dtau, Field<Real> ** fields, int nfields, double * params)
	{
		int kind = -1;
		if (fn == &update_pos) kind = GEVB_UPDATE_Q; else if (fn == &update_pos_Newton) kind = GEVB_UPDATE_Q_NEWTON; else if (fn == &displace_pcls_ic_basic) kind = GEVB_DISPLACE_PCLS_IC_BASIC;
		gevb_field * h[3] = {NULL, NULL, NULL};
		if (fields) handles(fields, nfields, h);
		check(gevb_moveParticles(p_, kind, dtau, h, fields ? nfields : 0, params), "Particles::moveParticles");
	}
	// with the callback's reduction output (ic_basic.hpp:1995: displace_pcls_ic_basic reports the largest displacement)
	void moveParticles(moveParticles_fn fn, double dtau, Field<Real> ** fields, int nfields, double * params, double * output, int * reduce_type, int noutput)
	{
		if (noutput <= 0 || output == NULL) { moveParticles(fn, dtau, fields, nfields, params); return; }
		int kind = (fn == &displace_pcls_ic_basic) ? GEVB_DISPLACE_PCLS_IC_BASIC : -1;
		if (noutput != 1 || reduce_type == NULL || reduce_type[0] != MAX) kind = -1;       // nothing else is used by the reference
		gevb_field * h[3] = {NULL, NULL, NULL};
		handles(fields, nfields, h);
		check(gevb_moveParticles_max(p_, kind, dtau, h, nfields, params, output), "Particles::moveParticles");
		lat_->max(output, 1);                                                            // LATfield2 reduces the outputs over the ranks
	}
	// fused form of main.cpp:775 + :798 (same result, one pass over the particles)
	Real kickDrift(updateVel_fn fn, double dtau_kick, int nf_kick, double * params_kick, double dtau_drift, int nf_drift, double * params_drift, Field<Real> ** fields)
	{
		int kind = (fn == &update_q) ? GEVB_UPDATE_Q : (fn == &update_q_Newton ? GEVB_UPDATE_Q_NEWTON : -1);
		gevb_field * h[3] = {NULL, NULL, NULL};
		handles(fields, 3, h);
		double maxvel = 0.;
		check(gevb_kick_drift(p_, kind, dtau_kick, nf_kick, params_kick, dtau_drift, nf_drift, params_drift, h, &maxvel), "Particles::kickDrift");
		return maxvel;
	}
};

// ---- projections (main.cpp:378-450) -------------------------------------------
inline void projection_init(Field<Real> * f) { check(gevb_projection_init(f->handle()), "projection_init"); }
inline void projection_T00_project(Particles_gevolution * pcls, Field<Real> * T00, double a = 1., Field<Real> * phi = NULL, double coeff = 1.)
{ check(gevb_projection_T00_project(pcls->handle(), T00->handle(), a, phi ? phi->handle() : NULL, coeff), "projection_T00_project"); }
inline void projection_T0i_project(Particles_gevolution * pcls, Field<Real> * T0i, Field<Real> * phi = NULL, double coeff = 1.)
{ check(gevb_projection_T0i_project(pcls->handle(), T0i->handle(), phi ? phi->handle() : NULL, coeff), "projection_T0i_project"); }
inline void projection_Tij_project(Particles_gevolution * pcls, Field<Real> * Tij, double a = 1., Field<Real> * phi = NULL, double coeff = 1.)
{ check(gevb_projection_Tij_project(pcls->handle(), Tij->handle(), a, phi ? phi->handle() : NULL, coeff), "projection_Tij_project"); }
inline void projection_T00_Tij_project(Particles_gevolution * pcls, Field<Real> * T00, Field<Real> * Tij, double a, Field<Real> * phi, double coeff = 1.)
{ check(gevb_projection_T00_Tij_project(pcls->handle(), T00->handle(), Tij->handle(), a, phi->handle(), coeff), "projection_T00_Tij_project"); }
inline void scalarProjectionCIC_project(Particles_gevolution * pcls, Field<Real> * rho) { check(gevb_scalarProjectionCIC_project(pcls->handle(), rho->handle()), "scalarProjectionCIC_project"); }
inline void scalarProjectionCIC_comm(Field<Real> * f) { check(gevb_projection_comm(f->handle()), "scalarProjectionCIC_comm"); }
inline void vectorProjectionCICNGP_comm(Field<Real> * f) { check(gevb_projection_comm(f->handle()), "vectorProjectionCICNGP_comm"); }
inline void symtensorProjectionCICNGP_comm(Field<Real> * f) { check(gevb_projection_comm(f->handle()), "symtensorProjectionCICNGP_comm"); }
#define projection_T00_comm scalarProjectionCIC_comm
#define projection_T0i_comm vectorProjectionCICNGP_comm
#define projection_Tij_comm symtensorProjectionCICNGP_comm

// ---- metric solve (gevolution.hpp:57-535) ----------------------------------------
template <class FieldType>
inline void prepareFTsource(Field<FieldType> & phi, Field<FieldType> & Tij, Field<FieldType> & Sij, const double coeff)
{ check(gevb_prepareFTsource_tensor(phi.handle(), Tij.handle(), Sij.handle(), coeff), "prepareFTsource"); }
template <class FieldType>
inline void prepareFTsource(Field<FieldType> & phi, Field<FieldType> & chi, Field<FieldType> & source, const FieldType bgmodel, Field<FieldType> & result, const double coeff, const double coeff2, const double coeff3)
{ check(gevb_prepareFTsource_scalar(phi.handle(), chi.handle(), source.handle(), bgmodel, result.handle(), coeff, coeff2, coeff3), "prepareFTsource"); }
inline void projectFTscalar(Field<Cplx> & SijFT, Field<Cplx> & chiFT, const int add = 0) { check(gevb_projectFTscalar(SijFT.handle(), chiFT.handle(), add), "projectFTscalar"); }
inline void evolveFTvector(Field<Cplx> & SijFT, Field<Cplx> & BiFT, const Real a2dtau) { check(gevb_evolveFTvector(SijFT.handle(), BiFT.handle(), a2dtau), "evolveFTvector"); }
// fused forms used by the time loop (same results, fewer passes over HBM)
inline void projectFTscalar_evolveFTvector(Field<Cplx> & SijFT, Field<Cplx> & chiFT, Field<Cplx> & BiFT, const Real a2dtau)
{ check(gevb_projectFTscalar_evolveFTvector(SijFT.handle(), chiFT.handle(), BiFT.handle(), a2dtau), "projectFTscalar_evolveFTvector"); }
template <class FieldType>
inline void prepareFTsource(Field<FieldType> & phi, Field<FieldType> & chi, Field<FieldType> & source, const FieldType bgmodel, Field<FieldType> & result, const double coeff, const double coeff2, const double coeff3, double & sum_source)
{ check(gevb_prepareFTsource_scalar_sum(phi.handle(), chi.handle(), source.handle(), bgmodel, result.handle(), coeff, coeff2, coeff3, &sum_source), "prepareFTsource"); }
// prepareFTsource + plan.execute(FFT_FORWARD) in one call (one pass over HBM where the own x-pass applies)
template <class T>
inline void prepareFTsource_execute(Field<Real> & phi, Field<Real> & chi, PlanFFT<T> & plan_source, const Real bgmodel, const double coeff, const double coeff2, const double coeff3, double * sum_source = NULL)
{ check(gevb_prepareFTsource_scalar_fft(phi.handle(), chi.handle(), plan_source.handle(), bgmodel, coeff, coeff2, coeff3, sum_source), "prepareFTsource"); }
template <class T>
inline void prepareFTsource_execute(Field<Real> & phi, PlanFFT<T> & plan_Sij, const double coeff)
{ check(gevb_prepareFTsource_tensor_fft(phi.handle(), plan_Sij.handle(), coeff), "prepareFTsource"); }
inline void projectFTvector(Field<Cplx> & SiFT, Field<Cplx> & BiFT, const Real coeff = 1., const Real modif = 0.) { check(gevb_projectFTvector(SiFT.handle(), BiFT.handle(), coeff, modif), "projectFTvector"); }
inline void projectFTtensor(Field<Cplx> & SijFT, Field<Cplx> & hijFT) { check(gevb_projectFTtensor(SijFT.handle(), hijFT.handle()), "projectFTtensor"); }
inline void solveModifiedPoissonFT(Field<Cplx> & sourceFT, Field<Cplx> & potFT, Real coeff, const Real modif = 0.) { check(gevb_solveModifiedPoissonFT(sourceFT.handle(), potFT.handle(), coeff, modif), "solveModifiedPoissonFT"); }

// ---- analysis (tools.hpp:237) --------------------------------------------------------
inline void extractPowerSpectrum(Field<Cplx> & fldFT, Real * kbin, Real * power, Real * kscatter, Real * pscatter, int * occupation, const int numbins, const bool deconvolve = true, const int ktype = 1)
{ check(gevb_extractPowerSpectrum(fldFT.handle(), kbin, power, kscatter, pscatter, occupation, numbins, deconvolve, ktype), "extractPowerSpectrum"); }

} // namespace gevb200

#endif
