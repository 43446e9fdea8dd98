// =============================================================================
// oracle/ref_driver.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE
// =============================================================================
// Builds oracle/_ref/libgevref.so: the REFERENCE's own hot-path source
// (/root/reference/gevolution.hpp, tools.hpp, background.hpp, metadata.hpp,
// #included by path at build time, never copied) compiled against the
// single-rank LATfield2 shim in oracle/latfield2_shim/, behind a flat-array
// C interface.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs may load this library; the product never does.
//
// Flat layouts (shared with oracle/gev_oracle.c and the tests):
//   real field    double[ncomp][N][N][N]          index [c][z][y][x], no halo
//   Fourier field double[ncomp][N][N][N/2+1][2]   index [c][kz][ky][kx][re,im]
//   particles     double pos[np][3], vel[np][3]   (vel = canonical momentum q/m)
//   symmetric tensor component order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
//
// The time-loop restatement in ref_sim_step() follows main.cpp:372-879 call by
// call (main.cpp itself cannot be compiled here: it needs HDF5 and the I/O
// modules).  The reference's own parser (parser.hpp), IC generator
// (ic_basic.hpp: generateIC_basic) and snapshot writer (Particles_gevolution.hpp:
// saveGadget2) ARE compiled in, over stand-ins for the GSL spline and MPI-IO
// calls they make (latfield2_shim/gsl/gsl_spline.h, LATfield2.hpp), so the shipped
// settings.ini can be run from its own seed: ref_sim_create_from_settings().
// =============================================================================
#include <stdint.h>
#include <stdlib.h>
#include <chrono>
#include <set>
#include <vector>
#include <unistd.h>
#include <fstream>
#include <iostream>
#include <string>
#include "LATfield2.hpp"
#include "metadata.hpp"
#include "class_tools.hpp"
#include "tools.hpp"
#include "background.hpp"
#include "Particles_gevolution.hpp"
#include "gevolution.hpp"
#include "ic_basic.hpp"
#include "parser.hpp"
#include "_ref/embedded_data.h"

using namespace std;
using namespace LATfield2;

typedef Particles_gevolution<part_simple, part_simple_info, part_simple_dataType> Pcls;

namespace {

struct Lat
{
	int N;
	Lattice lat, latFT, latPart;
	explicit Lat(int n) : N(n)
	{
		int box[3] = {n, n, n};
		lat.initialize(3, box, GRADIENT_ORDER);      // main.cpp:213
		latFT.initializeRealFFT(lat, 0);             // main.cpp:215
	}
};

Lat & get_lat(int N)
{
	static std::vector<Lat *> cache;
	for (size_t i = 0; i < cache.size(); i++) if (cache[i]->N == N) return *cache[i];
	cache.push_back(new Lat(N));
	return *cache.back();
}

void load_real(Field<Real> & f, const double * flat, bool halo = true)
{
	Lattice & l = f.lattice();
	const int N = l.size(0), nc = f.components();
	const size_t V = (size_t) N * N * N;
	for (int c = 0; c < nc; c++)
		for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
			f(l.indexOf(x, y, z), c) = flat[c * V + ((size_t) z * N + y) * N + x];
	if (halo) f.updateHalo();
}

void store_real(Field<Real> & f, double * flat)
{
	Lattice & l = f.lattice();
	const int N = l.size(0), nc = f.components();
	const size_t V = (size_t) N * N * N;
	for (int c = 0; c < nc; c++)
		for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
			flat[c * V + ((size_t) z * N + y) * N + x] = f(l.indexOf(x, y, z), c);
}

void load_cplx(Field<Cplx> & f, const double * flat)
{
	Lattice & l = f.lattice();
	const int nx = l.size(0), N = l.size(1), nc = f.components();
	const size_t Vk = (size_t) nx * N * N;
	for (int c = 0; c < nc; c++)
		for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < nx; x++)
		{
			size_t o = 2 * (c * Vk + ((size_t) z * N + y) * nx + x);
			f(l.indexOf(x, y, z), c) = Cplx(flat[o], flat[o + 1]);
		}
}

void store_cplx(Field<Cplx> & f, double * flat)
{
	Lattice & l = f.lattice();
	const int nx = l.size(0), N = l.size(1), nc = f.components();
	const size_t Vk = (size_t) nx * N * N;
	for (int c = 0; c < nc; c++)
		for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < nx; x++)
		{
			size_t o = 2 * (c * Vk + ((size_t) z * N + y) * nx + x);
			Cplx v = f(l.indexOf(x, y, z), c);
			flat[o] = v.real(); flat[o + 1] = v.imag();
		}
}

void make_pcls(Pcls & p, Lat & L, long np, const double * pos, const double * vel, double mass, const int64_t * ids = NULL)
{
	part_simple_info info;
	info.mass = mass; info.relativistic = 0; strcpy(info.type_name, "part_simple");
	part_simple_dataType dt;
	Real box[3] = {1., 1., 1.};
	p.initialize(info, dt, &L.lat, box);
	for (long i = 0; i < np; i++)
	{
		part_simple q;
		q.ID = ids ? (long) ids[i] : i;
		for (int l = 0; l < 3; l++) { q.pos[l] = pos[3 * i + l]; q.vel[l] = vel ? vel[3 * i + l] : 0.; }
		p.addParticle_global(q);
	}
}

// write particles back to flat arrays at slot ID (IDs must be 0..np-1 for this)
void read_pcls_by_id(Pcls & p, double * pos, double * vel)
{
	Site x(p.lattice());
	for (x.first(); x.test(); x.next())
	{
		partList<part_simple> & cell = p.field()(x);
		for (std::list<part_simple>::iterator it = cell.parts.begin(); it != cell.parts.end(); ++it)
			for (int l = 0; l < 3; l++)
			{
				if (pos) pos[3 * (*it).ID + l] = (*it).pos[l];
				if (vel) vel[3 * (*it).ID + l] = (*it).vel[l];
			}
	}
}

} // namespace

extern "C" {

const char * ref_describe(void)
{
	return "reference gevolution.hpp/tools.hpp/background.hpp (gevolution 1.2) compiled with -DFFT3D -DPHINONLINEAR "
	       "against the single-rank LATfield2 shim (oracle/latfield2_shim)";
}

// ---- PlanFFT::execute (main.cpp:477,488,544,563,575,593) --------------------
void ref_fft_forward(int N, int ncomp, const double * real_in, double * cplx_out)
{
	Lat & L = get_lat(N);
	Field<Real> r; Field<Cplx> k;
	r.initialize(L.lat, ncomp); k.initialize(L.latFT, ncomp);
	PlanFFT<Cplx> plan(&r, &k);
	load_real(r, real_in, false);
	plan.execute(FFT_FORWARD);
	store_cplx(k, cplx_out);
}

void ref_fft_backward(int N, int ncomp, const double * cplx_in, double * real_out)
{
	Lat & L = get_lat(N);
	Field<Real> r; Field<Cplx> k;
	r.initialize(L.lat, ncomp); k.initialize(L.latFT, ncomp);
	PlanFFT<Cplx> plan(&r, &k);
	load_cplx(k, cplx_in);
	plan.execute(FFT_BACKWARD);
	store_real(r, real_out);
}

// ---- real-space source preparation (gevolution.hpp:57,170) ------------------
void ref_prepareFTsource_scalar(int N, const double * phi, const double * chi, const double * source, double bgmodel, double * result, double coeff, double coeff2, double coeff3)
{
	Lat & L = get_lat(N);
	Field<Real> fphi(L.lat, 1), fchi(L.lat, 1), fsrc(L.lat, 1);
	load_real(fphi, phi); load_real(fchi, chi); load_real(fsrc, source);
	prepareFTsource<Real>(fphi, fchi, fsrc, bgmodel, fsrc, coeff, coeff2, coeff3);   // aliasing as in main.cpp:472
	store_real(fsrc, result);
}

void ref_prepareFTsource_tensor(int N, const double * phi, const double * Tij, double * Sij, double coeff)
{
	Lat & L = get_lat(N);
	Field<Real> fphi(L.lat, 1), fT;
	fT.initialize(L.lat, 3, 3, symmetric); fT.alloc();
	load_real(fphi, phi); load_real(fT, Tij);
	prepareFTsource<Real>(fphi, fT, fT, coeff);                                     // aliasing as in main.cpp:539
	store_real(fT, Sij);
}

// ---- Fourier-space kernels (gevolution.hpp:211-535) ---------------------------
void ref_solveModifiedPoissonFT(int N, const double * src, double * pot, double coeff, double modif)
{
	Lat & L = get_lat(N);
	Field<Cplx> f(L.latFT, 1);
	load_cplx(f, src);
	solveModifiedPoissonFT(f, f, coeff, modif);                                      // in place as in main.cpp:483
	store_cplx(f, pot);
}

void ref_projectFTscalar(int N, const double * SijFT, double * chiFT, int add)
{
	Lat & L = get_lat(N);
	Field<Cplx> S, chi(L.latFT, 1);
	S.initialize(L.latFT, 3, 3, symmetric); S.alloc();
	load_cplx(S, SijFT);
	if (add) load_cplx(chi, chiFT);
	projectFTscalar(S, chi, add);
	store_cplx(chi, chiFT);
}

void ref_evolveFTvector(int N, const double * SijFT, double * BiFT, double a2dtau)
{
	Lat & L = get_lat(N);
	Field<Cplx> S, B(L.latFT, 3);
	S.initialize(L.latFT, 3, 3, symmetric); S.alloc();
	load_cplx(S, SijFT); load_cplx(B, BiFT);
	evolveFTvector(S, B, a2dtau);
	store_cplx(B, BiFT);
}

void ref_projectFTvector(int N, const double * SiFT, double * BiFT, double coeff, double modif)
{
	Lat & L = get_lat(N);
	Field<Cplx> B(L.latFT, 3);
	load_cplx(B, SiFT);
	projectFTvector(B, B, coeff, modif);                                             // in place as in main.cpp:580
	store_cplx(B, BiFT);
}

void ref_projectFTtensor(int N, const double * SijFT, double * hijFT)
{
	Lat & L = get_lat(N);
	Field<Cplx> S;
	S.initialize(L.latFT, 3, 3, symmetric); S.alloc();
	load_cplx(S, SijFT);
	projectFTtensor(S, S);                                                           // in place as in output.hpp:263
	store_cplx(S, hijFT);
}

// ---- particle -> mesh projections (+ the *_comm fold) --------------------------
void ref_projection_T00(int N, long np, const double * pos, const double * vel, double mass, double a, const double * phi, double coeff, double * T00)
{
	Lat & L = get_lat(N);
	Pcls p; make_pcls(p, L, np, pos, vel, mass);
	Field<Real> src(L.lat, 1), fphi(L.lat, 1);
	if (phi) load_real(fphi, phi);
	projection_init(&src);
	projection_T00_project(&p, &src, a, phi ? &fphi : (Field<Real> *) NULL, coeff);
	projection_T00_comm(&src);
	store_real(src, T00);
}

void ref_projection_T0i(int N, long np, const double * pos, const double * vel, double mass, const double * phi, double coeff, double * T0i)
{
	Lat & L = get_lat(N);
	Pcls p; make_pcls(p, L, np, pos, vel, mass);
	Field<Real> B(L.lat, 3), fphi(L.lat, 1);
	if (phi) load_real(fphi, phi);
	projection_init(&B);
	projection_T0i_project(&p, &B, phi ? &fphi : (Field<Real> *) NULL, coeff);
	projection_T0i_comm(&B);
	store_real(B, T0i);
}

void ref_projection_Tij(int N, long np, const double * pos, const double * vel, double mass, double a, const double * phi, double coeff, double * Tij)
{
	Lat & L = get_lat(N);
	Pcls p; make_pcls(p, L, np, pos, vel, mass);
	Field<Real> S, fphi(L.lat, 1);
	S.initialize(L.lat, 3, 3, symmetric); S.alloc();
	if (phi) load_real(fphi, phi);
	projection_init(&S);
	projection_Tij_project(&p, &S, a, phi ? &fphi : (Field<Real> *) NULL, coeff);
	projection_Tij_comm(&S);
	store_real(S, Tij);
}

void ref_scalarProjectionCIC(int N, long np, const double * pos, double mass, double * rho)
{
	Lat & L = get_lat(N);
	Pcls p; make_pcls(p, L, np, pos, NULL, mass);
	Field<Real> src(L.lat, 1);
	projection_init(&src);
	scalarProjectionCIC_project(&p, &src);
	scalarProjectionCIC_comm(&src);
	store_real(src, rho);
}

// ---- kick / drift (main.cpp:775,798; callbacks gevolution.hpp:570,709,810,900) --
// kind: 0 = update_q / update_pos (GR), 1 = *_Newton
double ref_updateVel(int N, long np, const double * pos, double * vel, int kind, double dtau, const double * phi, const double * chi, const double * Bi, int nfields, const double * params)
{
	Lat & L = get_lat(N);
	Pcls p; make_pcls(p, L, np, pos, vel, 1.0);
	Field<Real> fphi(L.lat, 1), fchi(L.lat, 1), fB(L.lat, 3);
	if (phi) load_real(fphi, phi);
	if (chi) load_real(fchi, chi);
	if (Bi) load_real(fB, Bi);
	Field<Real> * fields[3] = {&fphi, &fchi, &fB};
	double par[2] = {params[0], params[1]};
	double r = (kind == 0) ? p.updateVel(update_q, dtau, fields, nfields, par)
	                       : p.updateVel(update_q_Newton, dtau, fields, nfields, par);
	read_pcls_by_id(p, NULL, vel);
	return r;
}

void ref_moveParticles(int N, long np, double * pos, const double * vel, int kind, double dtau, const double * phi, const double * chi, const double * Bi, int nfields, const double * params)
{
	Lat & L = get_lat(N);
	Pcls p; make_pcls(p, L, np, pos, vel, 1.0);
	Field<Real> fphi(L.lat, 1), fchi(L.lat, 1), fB(L.lat, 3);
	if (phi) load_real(fphi, phi);
	if (chi) load_real(fchi, chi);
	if (Bi) load_real(fB, Bi);
	Field<Real> * fields[3] = {&fphi, &fchi, &fB};
	double par[2] = {params[0], params[1]};
	if (kind == 0) p.moveParticles(update_pos, dtau, fields, nfields, par);
	else p.moveParticles(update_pos_Newton, dtau, NULL, 0, par);
	read_pcls_by_id(p, pos, NULL);
}

// cell under which each particle is filed + per-cell counts (bit-exact contract)
void ref_cell_index(int N, long np, const double * pos, int32_t * cell, uint32_t * counts)
{
	Lat & L = get_lat(N);
	Pcls p; make_pcls(p, L, np, pos, NULL, 1.0);
	if (counts) memset(counts, 0, sizeof(uint32_t) * (size_t) N * N * N);
	Site x(p.lattice());
	for (x.first(); x.test(); x.next())
	{
		partList<part_simple> & c = p.field()(x);
		int32_t key = (x.coord(2) * N + x.coord(1)) * N + x.coord(0);
		if (counts) counts[key] = (uint32_t) c.size;
		if (cell) for (std::list<part_simple>::iterator it = c.parts.begin(); it != c.parts.end(); ++it) cell[(*it).ID] = key;
	}
}

// ---- analysis (tools.hpp:53,364,405) -------------------------------------------
void ref_extractPowerSpectrum(int N, int ncomp, int symm, const double * fldFT, double * kbin, double * power, double * kscatter, double * pscatter, int * occupation, int numbins, int deconvolve, int ktype)
{
	Lat & L = get_lat(N);
	Field<Cplx> f;
	if (symm) f.initialize(L.latFT, 3, 3, symmetric); else f.initialize(L.latFT, ncomp);
	f.alloc();
	load_cplx(f, fldFT);
	extractPowerSpectrum(f, kbin, power, kscatter, pscatter, occupation, numbins, deconvolve != 0, ktype);
}

void ref_computeVectorDiagnostics(int N, const double * Bi, double * mdivB, double * mcurlB)
{
	Lat & L = get_lat(N);
	Field<Real> B(L.lat, 3);
	load_real(B, Bi);
	computeVectorDiagnostics(B, *mdivB, *mcurlB);
}

void ref_computeTensorDiagnostics(int N, const double * hij, double * mdivh, double * mtraceh, double * mnormh)
{
	Lat & L = get_lat(N);
	Field<Real> h;
	h.initialize(L.lat, 3, 3, symmetric); h.alloc();
	load_real(h, hij);
	computeTensorDiagnostics(h, *mdivh, *mtraceh, *mnormh);
}

// ---- background (background.hpp:137,167,200) ----------------------------------
// cosmo_in: Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, w0_fld, wa_fld, Omega_g, Omega_ur, Omega_rad, h
static cosmology make_cosmo(const double * c)
{
	cosmology co;
	memset(&co, 0, sizeof(co));
	co.Omega_cdm = c[0]; co.Omega_b = c[1]; co.Omega_m = c[2]; co.Omega_Lambda = c[3];
	co.Omega_fld = c[4]; co.w0_fld = c[5]; co.wa_fld = c[6]; co.Omega_g = c[7]; co.Omega_ur = c[8];
	co.Omega_rad = c[9]; co.h = c[10]; co.cs2_fld = 1.; co.num_ncdm = 0;
	return co;
}
double ref_Hconf(double a, double fourpiG, const double * cosmo) { return Hconf(a, fourpiG, make_cosmo(cosmo)); }
double ref_rungekutta4bg(double a, double fourpiG, const double * cosmo, double dtau) { rungekutta4bg(a, fourpiG, make_cosmo(cosmo), dtau); return a; }
double ref_particleHorizon(double a, double fourpiG, const double * cosmo) { cosmology co = make_cosmo(cosmo); return particleHorizon(a, fourpiG, co); }

// =============================================================================
// Stateful single-rank simulation: state of main.cpp:217-246 + the time loop
// =============================================================================
struct RefSim
{
	int N;
	Lat * L;
	cosmology cosmo;
	double boxsize, Cf, steplimit, z_in, z_relax;
	int gr_flag, vector_flag, baryon_flag;
	double fourpiG, a, tau, dtau, dtau_old, dx, T00hom;
	int cycle;
	double maxvel[2 + MAX_PCL_SPECIES - 2];    // by species slot: cdm, baryons, ncdm i (main.cpp indexes [i+1+baryon_flag])
	Pcls pcls_cdm, pcls_b;
	Pcls pcls_ncdm[MAX_PCL_SPECIES - 2];       // main.cpp:219
	bool have_ncdm[MAX_PCL_SPECIES - 2];
	double z_switch_deltancdm[MAX_PCL_SPECIES - 2], z_switch_Bncdm[MAX_PCL_SPECIES - 2], z_switch_linearchi, movelimit;
	int numsteps_ncdm[MAX_PCL_SPECIES - 2];
	bool have_b;
	Field<Real> phi, source, chi, Sij, Bi;
	Field<Cplx> scalarFT, SijFT, BiFT;
	PlanFFT<Cplx> plan_source, plan_phi, plan_chi, plan_Sij, plan_Bi;
	// BENCHMARK-style timers (main.cpp:71-88)
	double projection_time, gravity_solver_time, fft_time, update_q_time, moveParts_time, cycle_time;
};

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// settings: N, gr_flag, vector_flag(0 parabolic,1 elliptic) ; dsettings: boxsize, Cf, steplimit, z_in, z_relax
void * ref_sim_create(int N, int gr_flag, int vector_flag, const double * dsettings, const double * cosmo)
{
	RefSim * s = new RefSim();
	s->N = N; s->L = &get_lat(N);
	s->cosmo = make_cosmo(cosmo);
	bg_ncdm(-1., s->cosmo);   // flush background.hpp's per-process cache of the last bg_ncdm value (an earlier model may have had ncdm)
	s->boxsize = dsettings[0]; s->Cf = dsettings[1]; s->steplimit = dsettings[2]; s->z_in = dsettings[3]; s->z_relax = dsettings[4];
	s->gr_flag = gr_flag; s->vector_flag = vector_flag; s->baryon_flag = 0; s->have_b = false;
	Lattice & lat = s->L->lat; Lattice & latFT = s->L->latFT;
	// main.cpp:234-246
	s->source.initialize(lat, 1); s->phi.initialize(lat, 1); s->chi.initialize(lat, 1);
	s->scalarFT.initialize(latFT, 1);
	s->plan_source.initialize(&s->source, &s->scalarFT);
	s->plan_phi.initialize(&s->phi, &s->scalarFT);
	s->plan_chi.initialize(&s->chi, &s->scalarFT);
	s->Sij.initialize(lat, 3, 3, symmetric); s->SijFT.initialize(latFT, 3, 3, symmetric);
	s->plan_Sij.initialize(&s->Sij, &s->SijFT);
	s->Bi.initialize(lat, 3); s->BiFT.initialize(latFT, 3);
	s->plan_Bi.initialize(&s->Bi, &s->BiFT);
	// main.cpp:278-297
	s->dx = 1.0 / (double) N;
	s->fourpiG = 1.5 * s->boxsize * s->boxsize / C_SPEED_OF_LIGHT / C_SPEED_OF_LIGHT;
	s->a = 1. / (1. + s->z_in);
	s->tau = particleHorizon(s->a, s->fourpiG, s->cosmo);
	if (s->Cf * s->dx < s->steplimit / Hconf(s->a, s->fourpiG, s->cosmo)) s->dtau = s->Cf * s->dx;
	else s->dtau = s->steplimit / Hconf(s->a, s->fourpiG, s->cosmo);
	s->dtau_old = 0.;
	s->cycle = 0; s->T00hom = 0.;
	for (int i = 0; i < MAX_PCL_SPECIES; i++) s->maxvel[i] = 0.;
	for (int i = 0; i < MAX_PCL_SPECIES - 2; i++) { s->have_ncdm[i] = false; s->z_switch_deltancdm[i] = s->z_switch_Bncdm[i] = 0.; s->numsteps_ncdm[i] = 1; }
	s->z_switch_linearchi = 0.; s->movelimit = 1.e10;
	s->projection_time = s->gravity_solver_time = s->fft_time = s->update_q_time = s->moveParts_time = s->cycle_time = 0.;
	return s;
}

void ref_sim_destroy(void * h) { delete (RefSim *) h; }

// ---- the shipped configuration from its own seed -------------------------------------------------------------------
// settings.ini, the transfer-function table and the particle template of the reference (embedded at build time) are
// written to a scratch directory, parsed by the reference's parser (main.cpp:184-186), Ngrid / tiling factor / seed are
// overridden on request, and the reference's generateIC_basic (ic_basic.hpp:1626; call main.cpp:300) fills particles
// and metric fields.  Loop scalars as main.cpp:278-297,325-333.
static std::string write_embedded(const std::string & dir, const char * name, const unsigned char * data, unsigned long len)
{
	std::string path = dir + "/" + name;
	std::ofstream f(path.c_str(), std::ios::binary);
	f.write((const char *) data, (std::streamsize) len);
	return path;
}

// settings text with `overrides` applied: every line "key = value" of overrides replaces the line of the same key
// (or is appended), so the tests can switch e.g. "gravity theory = Newton" or "vector method = elliptic"
static std::string apply_overrides(const std::string & text, const char * overrides)
{
	if (overrides == NULL || overrides[0] == 0) return text;
	auto key_of = [](const std::string & line) { size_t e = line.find('='); if (e == std::string::npos) return std::string(); std::string k = line.substr(0, e);
		while (!k.empty() && (k.back() == ' ' || k.back() == '\t')) k.pop_back(); size_t b = k.find_first_not_of(" \t"); return b == std::string::npos ? std::string() : k.substr(b); };
	std::vector<std::string> lines, extra;
	std::string cur;
	for (char ch : text) { if (ch == '\n') { lines.push_back(cur); cur.clear(); } else cur += ch; }
	if (!cur.empty()) lines.push_back(cur);
	cur.clear();
	for (const char * q = overrides; ; q++) { if (*q == '\n' || *q == 0) { if (!cur.empty()) extra.push_back(cur); cur.clear(); if (*q == 0) break; } else cur += *q; }
	for (const std::string & o : extra)
	{
		const std::string k = key_of(o);
		bool done = false;
		for (std::string & l : lines) if (!k.empty() && l[0] != '#' && key_of(l) == k) { l = o; done = true; break; }
		if (!done) lines.push_back(o);
	}
	std::string out;
	for (const std::string & l : lines) out += l + "\n";
	return out;
}

// the reference's own parser (loadParameterFile + parseMetadata, parser.hpp:122,759) on a settings text: what the product's
// gevb_settings_read is compared with.  ints[16]: numpts, gr_flag, vector_flag, baryon_flag, seed, ksphere, correct_displacement,
// numtile[0], numtile[1], tracer_factor[0], tracer_factor[1], numbins, out_pk, out_snapshot, num_pk, num_snapshot;
// dbl[9 + 11 + 2 x 32]: boxsize, Cf, steplimit, movelimit, z_in, z_relax, A_s, n_s, k_pivot, the cosmology in gevb_settings' order,
// z_pk[32], z_snapshot[32].  Returns the number of parameters read (<= 0: failure).
int ref_parse_settings(const char * text, int * ints, double * dbl)
{
	char tmpl[] = "/tmp/gevref_XXXXXX";
	if (mkdtemp(tmpl) == NULL) return -1;
	const std::string dir(tmpl);
	const std::string settings = write_embedded(dir, "settings.ini", (const unsigned char *) text, strlen(text));
	metadata sim;
	icsettings ic;
	cosmology cosmo;
	parameter * params = NULL;
	const int numparam = loadParameterFile(settings.c_str(), params);
	if (numparam <= 0) return numparam;
	std::streambuf * quiet = std::cout.rdbuf(NULL);
	parseMetadata(params, numparam, sim, cosmo, ic);
	std::cout.rdbuf(quiet);
	free(params);
	remove(settings.c_str()); rmdir(dir.c_str());
	const int iv[16] = {sim.numpts, sim.gr_flag, sim.vector_flag, sim.baryon_flag, ic.seed, (ic.flags & ICFLAG_KSPHERE) ? 1 : 0, (ic.flags & ICFLAG_CORRECT_DISPLACEMENT) ? 1 : 0,
		ic.numtile[0], ic.numtile[1], sim.tracer_factor[0], sim.tracer_factor[1], sim.numbins, sim.out_pk, sim.out_snapshot, sim.num_pk, sim.num_snapshot};
	for (int i = 0; i < 16; i++) ints[i] = iv[i];
	const double dv[20] = {sim.boxsize, sim.Cf, sim.steplimit, sim.movelimit, sim.z_in, ic.z_relax, ic.A_s, ic.n_s, ic.k_pivot,
		cosmo.Omega_cdm, cosmo.Omega_b, cosmo.Omega_m, cosmo.Omega_Lambda, cosmo.Omega_fld, cosmo.w0_fld, cosmo.wa_fld, cosmo.Omega_g, cosmo.Omega_ur, cosmo.Omega_rad, cosmo.h};
	for (int i = 0; i < 20; i++) dbl[i] = dv[i];
	for (int i = 0; i < 32; i++) { dbl[20 + i] = i < sim.num_pk ? sim.z_pk[i] : 0.; dbl[52 + i] = i < sim.num_snapshot ? sim.z_snapshot[i] : 0.; }
	return numparam;
}

void * ref_sim_create_from_settings(int ngrid, int tiling, int seed, const char * overrides)
{
	char tmpl[] = "/tmp/gevref_XXXXXX";
	if (mkdtemp(tmpl) == NULL) return NULL;
	const std::string dir(tmpl);
	const std::string text = apply_overrides(std::string((const char *) ref_settings_ini, ref_settings_ini_len), overrides);
	const std::string settings = write_embedded(dir, "settings.ini", (const unsigned char *) text.data(), text.size());
	const std::string tkfile = write_embedded(dir, "class_tk.dat", ref_class_tk_dat, ref_class_tk_dat_len);
	const std::string pclfile = write_embedded(dir, "sc1_crystal.dat", ref_sc1_crystal_dat, ref_sc1_crystal_dat_len);
	metadata sim;
	icsettings ic;
	cosmology cosmo;
	parameter * params = NULL;
	int numparam = loadParameterFile(settings.c_str(), params);                       // main.cpp:184
	if (numparam <= 0) return NULL;
	std::streambuf * quiet = std::cout.rdbuf(NULL);                                  // the parser and the generator report on stdout
	parseMetadata(params, numparam, sim, cosmo, ic);                                 // main.cpp:186
	free(params); params = NULL; numparam = 0;
	if (ngrid > 0) sim.numpts = ngrid;
	if (tiling > 0) for (int i = 0; i < MAX_PCL_SPECIES; i++) if (ic.numtile[i] > 0 || i == 0) ic.numtile[i] = tiling;
	if (seed >= 0) ic.seed = seed;
	strcpy(ic.tkfile, tkfile.c_str());
	for (int i = 0; i < MAX_PCL_SPECIES; i++) strcpy(ic.pclfile[i], pclfile.c_str());
	const int N = sim.numpts;
	double ds[5] = {sim.boxsize, sim.Cf, sim.steplimit, sim.z_in, ic.z_relax};
	RefSim * s = new RefSim();
	s->N = N; s->L = &get_lat(N);
	s->cosmo = cosmo;
	bg_ncdm(-1., s->cosmo);
	s->boxsize = ds[0]; s->Cf = ds[1]; s->steplimit = ds[2]; s->z_in = ds[3]; s->z_relax = ds[4];
	s->gr_flag = sim.gr_flag; s->vector_flag = sim.vector_flag; s->baryon_flag = 0; s->have_b = false;
	Lattice & lat = s->L->lat; Lattice & latFT = s->L->latFT;
	s->source.initialize(lat, 1); s->phi.initialize(lat, 1); s->chi.initialize(lat, 1);
	s->scalarFT.initialize(latFT, 1);
	s->plan_source.initialize(&s->source, &s->scalarFT);
	s->plan_phi.initialize(&s->phi, &s->scalarFT);
	s->plan_chi.initialize(&s->chi, &s->scalarFT);
	s->Sij.initialize(lat, 3, 3, symmetric); s->SijFT.initialize(latFT, 3, 3, symmetric);
	s->plan_Sij.initialize(&s->Sij, &s->SijFT);
	s->Bi.initialize(lat, 3); s->BiFT.initialize(latFT, 3);
	s->plan_Bi.initialize(&s->Bi, &s->BiFT);
	for (int i = 0; i < MAX_PCL_SPECIES; i++) s->maxvel[i] = 0.;
	for (int i = 0; i < MAX_PCL_SPECIES - 2; i++) { s->have_ncdm[i] = false; s->z_switch_deltancdm[i] = sim.z_switch_deltancdm[i]; s->z_switch_Bncdm[i] = sim.z_switch_Bncdm[i]; s->numsteps_ncdm[i] = 1; }
	s->z_switch_linearchi = sim.z_switch_linearchi; s->movelimit = sim.movelimit;
	s->dx = 1.0 / (double) N;
	s->fourpiG = 1.5 * sim.boxsize * sim.boxsize / C_SPEED_OF_LIGHT / C_SPEED_OF_LIGHT;                      // main.cpp:280
	s->a = 1. / (1. + sim.z_in);
	s->tau = particleHorizon(s->a, s->fourpiG, s->cosmo);
	if (sim.Cf * s->dx < sim.steplimit / Hconf(s->a, s->fourpiG, s->cosmo)) s->dtau = sim.Cf * s->dx;          // main.cpp:292-295
	else s->dtau = sim.steplimit / Hconf(s->a, s->fourpiG, s->cosmo);
	s->dtau_old = 0.;
	s->cycle = 0; s->T00hom = 0.;
	s->projection_time = s->gravity_solver_time = s->fft_time = s->update_q_time = s->moveParts_time = s->cycle_time = 0.;
	double maxvel[MAX_PCL_SPECIES] = {0., 0., 0., 0., 0., 0.};
	generateIC_basic(sim, ic, s->cosmo, s->fourpiG, &s->pcls_cdm, &s->pcls_b, s->pcls_ncdm, maxvel, &s->phi, &s->chi, &s->Bi, &s->source, &s->Sij,
		&s->scalarFT, &s->BiFT, &s->SijFT, &s->plan_phi, &s->plan_chi, &s->plan_Bi, &s->plan_source, &s->plan_Sij, params, numparam);   // main.cpp:300
	std::cout.rdbuf(quiet);
	s->baryon_flag = sim.baryon_flag; s->have_b = sim.baryon_flag > 0;
	for (int i = 0; i < s->cosmo.num_ncdm; i++) s->have_ncdm[i] = sim.numpcl[1 + sim.baryon_flag + i] > 0;
	// main.cpp:325-333: maxvel[] in main.cpp's order (cdm, [baryons], ncdm...) -> species slots; gamma factor for GR
	const int numspecies = 1 + sim.baryon_flag + s->cosmo.num_ncdm;
	if (sim.gr_flag > 0) for (int i = 0; i < numspecies; i++) maxvel[i] /= sqrt(maxvel[i] * maxvel[i] + 1.0);
	s->maxvel[0] = maxvel[0];
	if (sim.baryon_flag) s->maxvel[1] = maxvel[1];
	for (int i = 0; i < s->cosmo.num_ncdm; i++) s->maxvel[2 + i] = maxvel[1 + sim.baryon_flag + i];
	std::remove(settings.c_str()); std::remove(tkfile.c_str()); std::remove(pclfile.c_str()); rmdir(dir.c_str());
	return s;
}

// ---- the IC generator's pieces by themselves (checkers of gevb_ic_cic_kernel / gevb_ic_displacement_field / gevb_ic_load_template)
// generateCICKernel (ic_basic.hpp:737): the whole kernel field, flat [z][y][x]
void ref_generateCICKernel(int N, long numpcl, float * pcldata, int numtile, double * out)
{
	Lat & L = get_lat(N);
	Field<Real> ker;
	ker.initialize(L.lat, 1); ker.alloc();
	if (numpcl > 0) generateCICKernel(ker, numpcl, pcldata, numtile); else generateCICKernel(ker);
	store_real(ker, out);
}

// generateDisplacementField (ic_basic.hpp:1090) on a Fourier field given and returned as flat [kz][ky][kx][2]; the spline is
// built from (x, y) with the shim's gsl_interp_cspline
void ref_generateDisplacementField(int N, double * potFT, double coeff, int n, const double * x, const double * y, unsigned int seed, int ksphere, int deconvolve_f)
{
	Lat & L = get_lat(N);
	Field<Cplx> f;
	f.initialize(L.latFT, 1); f.alloc();
	load_cplx(f, potFT);
	gsl_spline * sp = gsl_spline_alloc(gsl_interp_cspline, n);
	gsl_spline_init(sp, x, y, n);
	generateDisplacementField(f, coeff, sp, seed, ksphere, deconvolve_f);
	gsl_spline_free(sp);
	store_cplx(f, potFT);
}

// loadHomogeneousTemplate (ic_basic.hpp:191) of the embedded sc1_crystal.dat; the files the shipped configuration reads are
// written to `dir` (settings.ini, class_tk.dat, sc1_crystal.dat) so that a product run can start from them on a box
// without the reference tree.  Returns the template's particle count; positions (box units) go to out[3 * cap].
long ref_dump_shipped_files(const char * dir, float * out, long cap)
{
	const std::string d(dir);
	write_embedded(d, "settings.ini", ref_settings_ini, ref_settings_ini_len);
	write_embedded(d, "class_tk.dat", ref_class_tk_dat, ref_class_tk_dat_len);
	const std::string pclfile = write_embedded(d, "sc1_crystal.dat", ref_sc1_crystal_dat, ref_sc1_crystal_dat_len);
	long numpart = 0;
	float * data = NULL;
	loadHomogeneousTemplate(pclfile.c_str(), numpart, data);
	for (long i = 0; i < 3 * numpart && i < 3 * cap; i++) out[i] = data[i];
	free(data);
	return numpart;
}

// the reference's own spectrum file writer (tools.hpp:268-346), EXACT_OUTPUT_REDSHIFTS branch included
void ref_writePowerSpectrum(const double * kbin, const double * power, const double * kscatter, const double * pscatter, const int * occupation, int numbins,
                            double rescalek, double rescalep, const char * filename, const char * description, double a, double z_target)
{
	std::vector<Real> k(kbin, kbin + numbins), p(power, power + numbins), ks(kscatter, kscatter + numbins), ps(pscatter, pscatter + numbins);
	std::vector<int> occ(occupation, occupation + numbins);
	writePowerSpectrum(k.data(), p.data(), ks.data(), ps.data(), occ.data(), numbins, rescalek, rescalep, filename, description, a, z_target);
}

// configuration of a simulation as the flat arrays the other constructors take
// cosmo11: Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, w0_fld, wa_fld, Omega_g, Omega_ur, Omega_rad, h
// ds5: boxsize, Cf, steplimit, z_in, z_relax ; flags4: Ngrid, gr_flag, vector_flag, baryon_flag ; mass[2]: particle masses cdm, baryons
void ref_sim_get_config(void * h, double * cosmo11, double * ds5, int * flags4, double * mass2)
{
	RefSim * s = (RefSim *) h;
	const cosmology & c = s->cosmo;
	const double co[11] = {c.Omega_cdm, c.Omega_b, c.Omega_m, c.Omega_Lambda, c.Omega_fld, c.w0_fld, c.wa_fld, c.Omega_g, c.Omega_ur, c.Omega_rad, c.h};
	for (int i = 0; i < 11; i++) cosmo11[i] = co[i];
	ds5[0] = s->boxsize; ds5[1] = s->Cf; ds5[2] = s->steplimit; ds5[3] = s->z_in; ds5[4] = s->z_relax;
	flags4[0] = s->N; flags4[1] = s->gr_flag; flags4[2] = s->vector_flag; flags4[3] = s->baryon_flag;
	mass2[0] = s->pcls_cdm.parts_info()->mass;
	mass2[1] = s->have_b ? s->pcls_b.parts_info()->mass : 0.;
}

// the reference's own Gadget-2 writer (Particles_gevolution.hpp:30-251) with the header of writeSnapshots
// (output.hpp:360-402, restated: output.hpp itself needs HDF5)
static void ref_save_gadget2_at(RefSim * s, int species, const char * filename, int tracer_factor, double dtau_pos, double dtau_vel, double time, double redshift);
int ref_sim_save_gadget2(void * h, int species, const char * filename, int tracer_factor, double dtau_pos, double dtau_vel)
{
	RefSim * s = (RefSim *) h;
	ref_save_gadget2_at(s, species, filename, tracer_factor, dtau_pos, dtau_vel, s->a, (1. / s->a) - 1.);   // output.hpp:390-391
	return 0;
}

static void ref_save_gadget2_at(RefSim * s, int species, const char * filename, int tracer_factor, double dtau_pos, double dtau_vel, double time, double redshift)
{
	Pcls & p = species == 0 ? s->pcls_cdm : (species == 1 ? s->pcls_b : s->pcls_ncdm[species - 2]);
	gadget2_header hdr;
	memset(&hdr, 0, sizeof(hdr));
	hdr.num_files = 1;
	hdr.Omega0 = s->cosmo.Omega_m; hdr.OmegaLambda = s->cosmo.Omega_Lambda; hdr.HubbleParam = s->cosmo.h;
	hdr.BoxSize = s->boxsize / GADGET_LENGTH_CONVERSION;
	hdr.time = time; hdr.redshift = redshift;
	const long n = p.numParticles();
	const long nsel = (n % tracer_factor) ? (1 + n / tracer_factor) : (n / tracer_factor);
	hdr.npart[1] = (uint32_t) (nsel % (1ll << 32)); hdr.npartTotal[1] = hdr.npart[1]; hdr.npartTotalHW[1] = (uint32_t) (nsel / (1ll << 32));
	hdr.mass[1] = (double) tracer_factor * C_RHO_CRIT * p.parts_info()->mass * s->boxsize * s->boxsize * s->boxsize / GADGET_MASS_CONVERSION;
	std::remove(filename);
	std::streambuf * quiet = std::cout.rdbuf(NULL);
	p.saveGadget2(std::string(filename), hdr, tracer_factor, dtau_pos, dtau_vel, &s->phi);
	std::cout.rdbuf(quiet);
}

// ncdm species: cosmo.*_ncdm (metadata.hpp:284-288) and the switches (metadata.hpp:224-238)
void ref_sim_set_ncdm(void * h, int num_ncdm, const double * m_ncdm, const double * T_ncdm, const double * Omega_ncdm,
                      const double * zsw_delta, const double * zsw_B, double z_switch_linearchi, double movelimit)
{
	RefSim * s = (RefSim *) h;
	s->cosmo.num_ncdm = num_ncdm;
	for (int i = 0; i < num_ncdm; i++)
	{
		s->cosmo.m_ncdm[i] = m_ncdm[i]; s->cosmo.T_ncdm[i] = T_ncdm[i]; s->cosmo.Omega_ncdm[i] = Omega_ncdm[i]; s->cosmo.deg_ncdm[i] = 1.;
		s->z_switch_deltancdm[i] = zsw_delta[i]; s->z_switch_Bncdm[i] = zsw_B[i];
	}
	s->z_switch_linearchi = z_switch_linearchi; s->movelimit = movelimit;
	bg_ncdm(-1., s->cosmo);   // background.hpp:105-117 caches the last value by scale factor alone: flush it, the model changed
	if (s->cycle == 0)
	{
		s->tau = particleHorizon(s->a, s->fourpiG, s->cosmo);
		if (s->Cf * s->dx < s->steplimit / Hconf(s->a, s->fourpiG, s->cosmo)) s->dtau = s->Cf * s->dx;
		else s->dtau = s->steplimit / Hconf(s->a, s->fourpiG, s->cosmo);
	}
}
void ref_sim_set_ncdm_maxvel(void * h, const double * maxvel) { RefSim * s = (RefSim *) h; for (int i = 0; i < MAX_PCL_SPECIES - 2; i++) s->maxvel[2 + i] = maxvel[i]; }
void ref_sim_get_ncdm_state(void * h, double * maxvel, int * numsteps)
{
	RefSim * s = (RefSim *) h;
	for (int i = 0; i < MAX_PCL_SPECIES - 2; i++) { maxvel[i] = s->maxvel[2 + i]; numsteps[i] = s->numsteps_ncdm[i]; }
}
double ref_bg_ncdm(double a, const double * cosmo, int num_ncdm, const double * m_ncdm, const double * T_ncdm, const double * Omega_ncdm)
{
	cosmology co = make_cosmo(cosmo);
	co.num_ncdm = num_ncdm;
	for (int i = 0; i < num_ncdm; i++) { co.m_ncdm[i] = m_ncdm[i]; co.T_ncdm[i] = T_ncdm[i]; co.Omega_ncdm[i] = Omega_ncdm[i]; }
	double r = 0.;
	for (int i = 0; i < num_ncdm; i++) r += bg_ncdm(a, co, i);
	return r;
}

void ref_sim_set_particles(void * h, int species, long np, const int64_t * ids, const double * pos, const double * vel, double mass)
{
	RefSim * s = (RefSim *) h;
	if (species == 0) make_pcls(s->pcls_cdm, *s->L, np, pos, vel, mass, ids);
	else if (species == 1) { make_pcls(s->pcls_b, *s->L, np, pos, vel, mass, ids); s->have_b = true; s->baryon_flag = 1; }
	else { make_pcls(s->pcls_ncdm[species - 2], *s->L, np, pos, vel, mass, ids); s->have_ncdm[species - 2] = true; }
}

// which: 0 phi, 1 chi, 2 Bi(3), 3 source, 4 Sij(6)  [real];  10 scalarFT, 11 BiFT(3), 12 SijFT(6) [Fourier]
static Field<Real> * real_field(RefSim * s, int which)
{
	switch (which) { case 0: return &s->phi; case 1: return &s->chi; case 2: return &s->Bi; case 3: return &s->source; case 4: return &s->Sij; }
	return NULL;
}
static Field<Cplx> * cplx_field(RefSim * s, int which)
{
	switch (which) { case 10: return &s->scalarFT; case 11: return &s->BiFT; case 12: return &s->SijFT; }
	return NULL;
}
void ref_sim_set_field(void * h, int which, const double * data)
{
	RefSim * s = (RefSim *) h;
	if (which < 10) load_real(*real_field(s, which), data); else load_cplx(*cplx_field(s, which), data);
}
void ref_sim_get_field(void * h, int which, double * data)
{
	RefSim * s = (RefSim *) h;
	if (which < 10) store_real(*real_field(s, which), data); else store_cplx(*cplx_field(s, which), data);
}
static Pcls & species_pcls(RefSim * s, int species) { return species == 0 ? s->pcls_cdm : (species == 1 ? s->pcls_b : s->pcls_ncdm[species - 2]); }
long ref_sim_num_particles(void * h, int species) { RefSim * s = (RefSim *) h; return species_pcls(s, species).numParticles(); }

// particles out in lattice iteration order (cell-sorted, x fastest)
void ref_sim_get_particles(void * h, int species, int64_t * ids, double * pos, double * vel)
{
	RefSim * s = (RefSim *) h;
	Pcls & p = species_pcls(s, species);
	Site x(p.lattice());
	long n = 0;
	for (x.first(); x.test(); x.next())
	{
		partList<part_simple> & cell = p.field()(x);
		for (std::list<part_simple>::iterator it = cell.parts.begin(); it != cell.parts.end(); ++it, ++n)
		{
			ids[n] = (*it).ID;
			for (int l = 0; l < 3; l++) { pos[3 * n + l] = (*it).pos[l]; vel[3 * n + l] = (*it).vel[l]; }
		}
	}
}

// scalars: a, tau, dtau, dtau_old, cycle, maxvel0, maxvel1, T00hom, fourpiG
void ref_sim_get_state(void * h, double * out)
{
	RefSim * s = (RefSim *) h;
	out[0] = s->a; out[1] = s->tau; out[2] = s->dtau; out[3] = s->dtau_old; out[4] = s->cycle;
	out[5] = s->maxvel[0]; out[6] = s->maxvel[1]; out[7] = s->T00hom; out[8] = s->fourpiG;
}
void ref_sim_set_state(void * h, const double * in)
{
	RefSim * s = (RefSim *) h;
	s->a = in[0]; s->tau = in[1]; s->dtau = in[2]; s->dtau_old = in[3]; s->cycle = (int) in[4];
	s->maxvel[0] = in[5]; s->maxvel[1] = in[6];
}
// timers: projection, gravity solver, thereof FFT, update momenta, move particles, cycle total
void ref_sim_get_timers(void * h, double * out)
{
	RefSim * s = (RefSim *) h;
	out[0] = s->projection_time; out[1] = s->gravity_solver_time; out[2] = s->fft_time;
	out[3] = s->update_q_time; out[4] = s->moveParts_time; out[5] = s->cycle_time;
}

// one cycle of the main loop, outputs stripped (main.cpp:372-879)
static void ref_sim_solve(RefSim * s);
static void ref_sim_update(RefSim * s);
void ref_sim_step(void * h) { ref_sim_solve((RefSim *) h); ref_sim_update((RefSim *) h); }

// first half of the cycle: projections and metric solve (main.cpp:378-599)
static void ref_sim_solve(RefSim * s)
{
	const double dx = s->dx, fourpiG = s->fourpiG;
	cosmology & cosmo = s->cosmo;
	double & a = s->a; double & dtau_old = s->dtau_old;
	Site x(s->L->lat);
	double t0 = now_s(), t1, t2;

	// main.cpp:378-411  T00
	projection_init(&s->source);
	if (s->gr_flag > 0)
	{
		projection_T00_project(&s->pcls_cdm, &s->source, a, &s->phi);
		if (s->baryon_flag) projection_T00_project(&s->pcls_b, &s->source, a, &s->phi);
		for (int i = 0; i < cosmo.num_ncdm; i++)     // main.cpp:388-399 (radiation_flag == 0)
		{
			if (a >= 1. / (s->z_switch_deltancdm[i] + 1.) && s->have_ncdm[i])
				projection_T00_project(s->pcls_ncdm + i, &s->source, a, &s->phi);
			else
			{
				double tmp = bg_ncdm(a, cosmo, i);
				for (x.first(); x.test(); x.next()) s->source(x) += tmp;
			}
		}
	}
	else
	{
		scalarProjectionCIC_project(&s->pcls_cdm, &s->source);
		if (s->baryon_flag) scalarProjectionCIC_project(&s->pcls_b, &s->source);
		for (int i = 0; i < cosmo.num_ncdm; i++)     // main.cpp:405-409
			if (a >= 1. / (s->z_switch_deltancdm[i] + 1.) && s->have_ncdm[i]) scalarProjectionCIC_project(s->pcls_ncdm + i, &s->source);
	}
	projection_T00_comm(&s->source);

	// main.cpp:424-436  T0i (elliptic only)
	if (s->vector_flag == VECTOR_ELLIPTIC)
	{
		projection_init(&s->Bi);
		projection_T0i_project(&s->pcls_cdm, &s->Bi, &s->phi);
		if (s->baryon_flag) projection_T0i_project(&s->pcls_b, &s->Bi, &s->phi);
		for (int i = 0; i < cosmo.num_ncdm; i++)     // main.cpp:430-434
			if (a >= 1. / (s->z_switch_Bncdm[i] + 1.) && s->have_ncdm[i]) projection_T0i_project(s->pcls_ncdm + i, &s->Bi, &s->phi);
		projection_T0i_comm(&s->Bi);
	}

	// main.cpp:438-450  Tij
	projection_init(&s->Sij);
	projection_Tij_project(&s->pcls_cdm, &s->Sij, a, &s->phi);
	if (s->baryon_flag) projection_Tij_project(&s->pcls_b, &s->Sij, a, &s->phi);
	if (a >= 1. / (s->z_switch_linearchi + 1.))      // main.cpp:442-449
		for (int i = 0; i < cosmo.num_ncdm; i++)
			if (s->have_ncdm[i]) projection_Tij_project(s->pcls_ncdm + i, &s->Sij, a, &s->phi);
	projection_Tij_comm(&s->Sij);

	t1 = now_s(); s->projection_time += t1 - t0;

	if (s->gr_flag > 0)
	{
		// main.cpp:459-463
		double T00hom = 0.;
		for (x.first(); x.test(); x.next()) T00hom += s->source(x);
		T00hom /= (double) ((long) s->N * s->N * s->N);
		s->T00hom = T00hom;

		if (dtau_old > 0.)
		{
			// main.cpp:472-488
			prepareFTsource<Real>(s->phi, s->chi, s->source, cosmo.Omega_cdm + cosmo.Omega_b + bg_ncdm(a, cosmo), s->source, 3. * Hconf(a, fourpiG, cosmo) * dx * dx / dtau_old, fourpiG * dx * dx / a, 3. * Hconf(a, fourpiG, cosmo) * Hconf(a, fourpiG, cosmo) * dx * dx);
			t2 = now_s(); s->plan_source.execute(FFT_FORWARD); s->fft_time += now_s() - t2;
			solveModifiedPoissonFT(s->scalarFT, s->scalarFT, 1. / (dx * dx), 3. * Hconf(a, fourpiG, cosmo) / dtau_old);
			t2 = now_s(); s->plan_phi.execute(FFT_BACKWARD); s->fft_time += now_s() - t2;
		}
	}
	else
	{
		// main.cpp:500-511
		t2 = now_s(); s->plan_source.execute(FFT_FORWARD); s->fft_time += now_s() - t2;
		solveModifiedPoissonFT(s->scalarFT, s->scalarFT, fourpiG / a);
		t2 = now_s(); s->plan_phi.execute(FFT_BACKWARD); s->fft_time += now_s() - t2;
	}

	s->phi.updateHalo();   // main.cpp:518

	// main.cpp:539-568  chi
	prepareFTsource<Real>(s->phi, s->Sij, s->Sij, 2. * fourpiG * dx * dx / a);
	t2 = now_s(); s->plan_Sij.execute(FFT_FORWARD); s->fft_time += now_s() - t2;
	projectFTscalar(s->SijFT, s->scalarFT);
	t2 = now_s(); s->plan_chi.execute(FFT_BACKWARD); s->fft_time += now_s() - t2;
	s->chi.updateHalo();

	// main.cpp:570-599  B
	if (s->vector_flag == VECTOR_ELLIPTIC)
	{
		t2 = now_s(); s->plan_Bi.execute(FFT_FORWARD); s->fft_time += now_s() - t2;
		projectFTvector(s->BiFT, s->BiFT, fourpiG * dx * dx);
	}
	else
		evolveFTvector(s->SijFT, s->BiFT, a * a * dtau_old);

	if (s->gr_flag > 0)
	{
		t2 = now_s(); s->plan_Bi.execute(FFT_BACKWARD); s->fft_time += now_s() - t2;
		s->Bi.updateHalo();
	}

	t2 = now_s(); s->gravity_solver_time += t2 - t1;
	s->cycle_time += now_s() - t0;
}

// second half of the cycle: particle updates, background, next time step (main.cpp:696-875)
static void ref_sim_update(RefSim * s)
{
	const double dx = s->dx, fourpiG = s->fourpiG;
	cosmology & cosmo = s->cosmo;
	double & a = s->a; double & dtau = s->dtau; double & dtau_old = s->dtau_old;
	Field<Real> * update_cdm_fields[3] = {&s->phi, &s->chi, &s->Bi};
	double f_params[5];
	double t0 = now_s(), t1, t2 = t0;

	// main.cpp:696-701  step subdivisions for the ncdm updates
	for (int i = 0; i < cosmo.num_ncdm; i++)
	{
		if (dtau * s->maxvel[2 + i] > dx * s->movelimit)
			s->numsteps_ncdm[i] = (int) ceil(dtau * s->maxvel[2 + i] / dx / s->movelimit);
		else s->numsteps_ncdm[i] = 1;
	}
	// main.cpp:729-765  non-cold DM particle update
	for (int i = 0; i < cosmo.num_ncdm; i++)
	{
		if (!s->have_ncdm[i]) continue;
		double tmp = a;
		for (int j = 0; j < s->numsteps_ncdm[i]; j++)
		{
			f_params[0] = tmp;
			f_params[1] = tmp * tmp * s->N;
			if (s->gr_flag > 0)
				s->maxvel[2 + i] = s->pcls_ncdm[i].updateVel(update_q, (dtau + dtau_old) / 2. / s->numsteps_ncdm[i], update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 2), f_params);
			else
				s->maxvel[2 + i] = s->pcls_ncdm[i].updateVel(update_q_Newton, (dtau + dtau_old) / 2. / s->numsteps_ncdm[i], update_cdm_fields, 1, f_params);
			rungekutta4bg(tmp, fourpiG, cosmo, 0.5 * dtau / s->numsteps_ncdm[i]);
			f_params[0] = tmp;
			f_params[1] = tmp * tmp * s->N;
			if (s->gr_flag > 0)
				s->pcls_ncdm[i].moveParticles(update_pos, dtau / s->numsteps_ncdm[i], update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 2), f_params);
			else
				s->pcls_ncdm[i].moveParticles(update_pos_Newton, dtau / s->numsteps_ncdm[i], NULL, 0, f_params);
			rungekutta4bg(tmp, fourpiG, cosmo, 0.5 * dtau / s->numsteps_ncdm[i]);
		}
	}

	// main.cpp:771-784  kick
	f_params[0] = a;
	f_params[1] = a * a * s->N;
	if (s->gr_flag > 0)
	{
		s->maxvel[0] = s->pcls_cdm.updateVel(update_q, (dtau + dtau_old) / 2., update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 2), f_params);
		if (s->baryon_flag) s->maxvel[1] = s->pcls_b.updateVel(update_q, (dtau + dtau_old) / 2., update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 2), f_params);
	}
	else
	{
		s->maxvel[0] = s->pcls_cdm.updateVel(update_q_Newton, (dtau + dtau_old) / 2., update_cdm_fields, 1, f_params);
		if (s->baryon_flag) s->maxvel[1] = s->pcls_b.updateVel(update_q_Newton, (dtau + dtau_old) / 2., update_cdm_fields, 1, f_params);
	}
	t1 = now_s(); s->update_q_time += t1 - t2;

	rungekutta4bg(a, fourpiG, cosmo, 0.5 * dtau);   // main.cpp:792

	// main.cpp:794-807  drift
	f_params[0] = a;
	f_params[1] = a * a * s->N;
	if (s->gr_flag > 0)
	{
		s->pcls_cdm.moveParticles(update_pos, dtau, update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 0), f_params);
		if (s->baryon_flag) s->pcls_b.moveParticles(update_pos, dtau, update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 0), f_params);
	}
	else
	{
		s->pcls_cdm.moveParticles(update_pos_Newton, dtau, NULL, 0, f_params);
		if (s->baryon_flag) s->pcls_b.moveParticles(update_pos_Newton, dtau, NULL, 0, f_params);
	}
	s->moveParts_time += now_s() - t1;

	rungekutta4bg(a, fourpiG, cosmo, 0.5 * dtau);   // main.cpp:814

	// main.cpp:816-822
	if (s->gr_flag > 0)
		for (int i = 0; i < 2 + cosmo.num_ncdm; i++) s->maxvel[i] /= sqrt(s->maxvel[i] * s->maxvel[i] + 1.0);

	s->tau += dtau;         // main.cpp:825
	dtau_old = dtau;        // main.cpp:867
	if (s->Cf * dx < s->steplimit / Hconf(a, fourpiG, cosmo)) dtau = s->Cf * dx;   // main.cpp:869-872
	else dtau = s->steplimit / Hconf(a, fourpiG, cosmo);
	s->cycle++;
	s->cycle_time += now_s() - t0;
}

// ---- the main loop with its outputs (main.cpp:605-693), as far as they can be produced without HDF5 ---------------------
// writeSpectra's phi / chi / hij / B branches (output.hpp:1945-1981,2151-2155; output.hpp itself needs HDF5, so the call
// sequence is restated here around the reference's own extractPowerSpectrum and writePowerSpectrum)
static void ref_write_spectra(RefSim * s, const char * prefix, int pkcount, int numbins, int mask, double z_target)
{
	const double a = s->a, fourpiG = s->fourpiG;
	const Real numpts3d = (Real) s->N * (Real) s->N * (Real) s->N;
	std::vector<Real> kbin(numbins), power(numbins), kscatter(numbins), pscatter(numbins);
	std::vector<int> occupation(numbins);
	char filename[1024];
	if (mask & MASK_PHI)
	{
		s->plan_phi.execute(FFT_FORWARD);
		extractPowerSpectrum(s->scalarFT, kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, false, KTYPE_LINEAR);
		sprintf(filename, "%s%03d_phi.dat", prefix, pkcount);
		writePowerSpectrum(kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, s->boxsize, (Real) numpts3d * (Real) numpts3d * 2. * M_PI * M_PI, filename, "power spectrum of phi", a, z_target);
	}
	if (mask & MASK_CHI)
	{
		s->plan_chi.execute(FFT_FORWARD);
		extractPowerSpectrum(s->scalarFT, kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, false, KTYPE_LINEAR);
		sprintf(filename, "%s%03d_chi.dat", prefix, pkcount);
		writePowerSpectrum(kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, s->boxsize, (Real) numpts3d * (Real) numpts3d * 2. * M_PI * M_PI, filename, "power spectrum of chi", a, z_target);
	}
	if (mask & MASK_HIJ)
	{
		projection_init(&s->Sij);
		projection_Tij_project(&s->pcls_cdm, &s->Sij, a, &s->phi);
		if (s->baryon_flag) projection_Tij_project(&s->pcls_b, &s->Sij, a, &s->phi);
		for (int i = 0; i < s->cosmo.num_ncdm; i++) if (s->have_ncdm[i]) projection_Tij_project(s->pcls_ncdm + i, &s->Sij, a, &s->phi);
		projection_Tij_comm(&s->Sij);
		prepareFTsource<Real>(s->phi, s->Sij, s->Sij, 2. * fourpiG / (double) s->N / (double) s->N / a);
		s->plan_Sij.execute(FFT_FORWARD);
		projectFTtensor(s->SijFT, s->SijFT);
		extractPowerSpectrum(s->SijFT, kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, false, KTYPE_LINEAR);
		sprintf(filename, "%s%03d_hij.dat", prefix, pkcount);
		writePowerSpectrum(kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, s->boxsize, 2. * M_PI * M_PI, filename, "power spectrum of hij", a, z_target);
	}
	if (mask & MASK_B)
	{
		extractPowerSpectrum(s->BiFT, kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, false, KTYPE_LINEAR);
		sprintf(filename, "%s%03d_B.dat", prefix, pkcount);
		writePowerSpectrum(kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, s->boxsize, a * a * a * a * s->N * s->N * 2. * M_PI * M_PI, filename, "power spectrum of B", a, z_target);
	}
}

int ref_sim_run(void * h, const double * z_pk, int num_pk, int pk_mask, int numbins, const char * pk_prefix,
                const double * z_snapshot, int num_snapshot, int tracer_factor, const char * snap_prefix, int max_cycles, int * counts3)
{
	RefSim * s = (RefSim *) h;
	int pkcount = 0, snapcount = 0, cycles = 0;
	std::streambuf * quiet = std::cout.rdbuf(NULL);
	while (cycles < max_cycles)
	{
		ref_sim_solve(s);
		const double a = s->a;
		if (snapcount < num_snapshot && 1. / a < z_snapshot[snapcount] + 1.)                     // main.cpp:617
		{
			const double time = 1. / (z_snapshot[snapcount] + 1.);                               // output.hpp:386
			const double dtau_pos = (time - a) / a / Hconf(a, s->fourpiG, s->cosmo);             // output.hpp:388
			char name[1024];
			sprintf(name, "%s%03d_cdm", snap_prefix, snapcount);
			ref_save_gadget2_at(s, 0, name, tracer_factor, dtau_pos, dtau_pos + 0.5 * s->dtau_old, time, z_snapshot[snapcount]);   // output.hpp:409
			if (s->baryon_flag) { sprintf(name, "%s%03d_b", snap_prefix, snapcount); ref_save_gadget2_at(s, 1, name, tracer_factor, dtau_pos, dtau_pos + 0.5 * s->dtau_old, time, z_snapshot[snapcount]); }
			snapcount++;
		}
		if (pkcount < num_pk && 1. / a < z_pk[pkcount] + 1.)                                     // main.cpp:641
		{
			ref_write_spectra(s, pk_prefix, pkcount, numbins, pk_mask, z_pk[pkcount]);
			pkcount++;
		}
		double tmp = a;                                                                          // main.cpp:661-664
		rungekutta4bg(tmp, s->fourpiG, s->cosmo, 0.5 * s->dtau);
		rungekutta4bg(tmp, s->fourpiG, s->cosmo, 0.5 * s->dtau);
		if (pkcount < num_pk && 1. / tmp < z_pk[pkcount] + 1.)                                   // main.cpp:666
			ref_write_spectra(s, pk_prefix, pkcount, numbins, pk_mask, z_pk[pkcount]);
		if (pkcount >= num_pk && snapcount >= num_snapshot) break;                               // main.cpp:685-693
		ref_sim_update(s);
		cycles++;
	}
	std::cout.rdbuf(quiet);
	if (counts3) { counts3[0] = cycles; counts3[1] = pkcount; counts3[2] = snapcount; }
	return 0;
}

} // extern "C"
