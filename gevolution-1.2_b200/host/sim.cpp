// sim.cpp -- the reference's time loop over the device-resident drop-in surface
//
// C++ host side of the hot path: one cycle of gevolution 1.2's main loop
// (main.cpp:372-879) with the output / hibernation branches stripped, written
// against include/gevolution_b200.hpp so that it reads like the reference's own
// loop.  Exposed through the C ABI as gevb_sim_* (include/gevb.h).
#include "sim_internal.hpp"

// No exception crosses the C boundary: every gevb_sim_* body that can allocate or call into the library runs inside
// GEVB_C_BOUNDARY, which turns gevb_error (message already in gevb_last_error) and anything else into status 1.
#define GEVB_C_BOUNDARY(body) try { body } catch (const gevb_error &) { return 1; } catch (...) { return 1; }

static int sim_create(gevb_sim * s, int gr_flag, int vector_flag, const double * ds, const double * c);

// main.cpp:281-286: no particle may move farther than the thinnest local domain minus one cell per update (the migration
// reaches the adjacent slab only); with z-slabs that is nz_local - 1 (N - 1 on one rank)
static double clamp_movelimit(gevb_sim * s, double movelimit)
{
	int rank = 0, nranks = 1, n = 0, nzl = 0;
	gevb_ctx_ranks(s->lat.ctx(), &rank, &nranks);
	gevb_ctx_geometry(s->lat.ctx(), &n, NULL, &nzl, NULL, NULL);
	const double lim = (double) ((nranks > 1 ? nzl : n) - 1);
	return movelimit < lim ? movelimit : lim;
}

extern "C" int gevb_sim_create(gevb_sim ** out, gevb_ctx * ctx, int gr_flag, int vector_flag, const double * ds, const double * c)
{
	if (out == NULL || ctx == NULL || ds == NULL || c == NULL) return 1;
	gevb_sim * s = NULL;
	// a failed field / plan allocation (the Sij and SijFT fields are 13 GB at 512^3) throws from the wrappers: the partly
	// built simulation is destroyed (the field destructors release what was allocated) and the status reports it
	try
	{
		s = new gevb_sim(ctx);
		if (sim_create(s, gr_flag, vector_flag, ds, c) != 0) { delete s; return 1; }
	}
	catch (...) { delete s; return 1; }
	*out = s;
	return 0;
}

static int sim_create(gevb_sim * s, int gr_flag, int vector_flag, const double * ds, const double * c)
{
	s->numpts = s->lat.size(0);
	s->gr_flag = gr_flag; s->vector_flag = vector_flag; s->baryon_flag = 0; s->fused = 1;
	s->boxsize = ds[0]; s->Cf = ds[1]; s->steplimit = ds[2]; s->z_in = ds[3]; s->z_relax = ds[4];
	cosmology co;
	std::memset(&co, 0, sizeof(co));
	co.Omega_cdm = c[0]; co.Omega_b = c[1]; co.Omega_m = c[2]; co.Omega_Lambda = c[3]; co.Omega_fld = c[4]; co.w0_fld = c[5]; co.wa_fld = c[6];
	co.Omega_g = c[7]; co.Omega_ur = c[8]; co.Omega_rad = c[9]; co.h = c[10]; co.num_ncdm = 0;
	s->cosmo = co;
	s->z_switch_linearchi = 0.;
	s->movelimit = clamp_movelimit(s, 1.e10);
	for (int i = 0; i < GEVB_MAX_NCDM; i++) { s->z_switch_deltancdm[i] = s->z_switch_Bncdm[i] = 0.; s->numsteps_ncdm[i] = 1; }
	Lattice & lat = s->lat;
	// main.cpp:234-246
	s->source.initialize(lat, 1);
	s->phi.initialize(lat, 1);
	s->chi.initialize(lat, 1);
	s->scalarFT.initialize(lat, 1);
	s->plan_source.initialize(&s->source, &s->scalarFT);
	s->plan_phi.initialize(&s->phi, &s->scalarFT);
	s->plan_chi.initialize(&s->chi, &s->scalarFT);
	// scalarFT is scratch: every backward transform of it is followed by a forward one that overwrites it
	s->plan_phi.preserveInput(false);
	s->plan_chi.preserveInput(false);
	s->Sij.initialize(lat, 3, 3, symmetric);
	s->SijFT.initialize(lat, 3, 3, symmetric);
	s->plan_Sij.initialize(&s->Sij, &s->SijFT);
	s->Bi.initialize(lat, 3);
	s->BiFT.initialize(lat, 3);
	s->plan_Bi.initialize(&s->Bi, &s->BiFT);
	// main.cpp:278-297
	s->dx = 1.0 / (double) s->numpts;
	s->fourpiG = 1.5 * s->boxsize * s->boxsize / GEVB_C_SPEED_OF_LIGHT / GEVB_C_SPEED_OF_LIGHT;
	s->a = 1. / (1. + s->z_in);
	s->tau = particleHorizon(s->a, s->fourpiG, s->cosmo);
	if (s->Cf * s->dx < s->steplimit / Hconf(s->a, s->fourpiG, s->cosmo)) s->dtau = s->Cf * s->dx;
	else s->dtau = s->steplimit / Hconf(s->a, s->fourpiG, s->cosmo);
	s->dtau_old = 0.;
	s->cycle = 0; s->T00hom = 0.;
	for (int i = 0; i < 2 + GEVB_MAX_NCDM; i++) s->maxvel[i] = 0.;
	return 0;
}

// non-cold dark matter species (cosmo.*_ncdm, metadata.hpp:284-288; switches metadata.hpp:224-238, parser.hpp:1755-1793)
extern "C" int gevb_sim_set_ncdm(gevb_sim * s, int num_ncdm, const double * m_ncdm, const double * T_ncdm, const double * Omega_ncdm,
                                 const double * z_switch_deltancdm, const double * z_switch_Bncdm, double z_switch_linearchi, double movelimit)
{
	if (s == NULL || num_ncdm < 0 || num_ncdm > GEVB_MAX_NCDM) return 1;
	if (num_ncdm > 0 && (m_ncdm == NULL || T_ncdm == NULL || Omega_ncdm == NULL || z_switch_deltancdm == NULL || z_switch_Bncdm == NULL)) return 1;
	s->cosmo.num_ncdm = num_ncdm;
	for (int i = 0; i < num_ncdm; i++)
	{
		s->cosmo.m_ncdm[i] = m_ncdm[i]; s->cosmo.T_ncdm[i] = T_ncdm[i]; s->cosmo.Omega_ncdm[i] = Omega_ncdm[i];
		s->z_switch_deltancdm[i] = z_switch_deltancdm[i]; s->z_switch_Bncdm[i] = z_switch_Bncdm[i];
	}
	s->z_switch_linearchi = z_switch_linearchi; s->movelimit = clamp_movelimit(s, movelimit);
	if (s->cycle == 0)
	{
		// the background changed: redo main.cpp:289-295 at the initial redshift
		s->tau = particleHorizon(s->a, s->fourpiG, s->cosmo);
		if (s->Cf * s->dx < s->steplimit / Hconf(s->a, s->fourpiG, s->cosmo)) s->dtau = s->Cf * s->dx;
		else s->dtau = s->steplimit / Hconf(s->a, s->fourpiG, s->cosmo);
	}
	return 0;
}

extern "C" int gevb_sim_get_ncdm_state(gevb_sim * s, double * maxvel, int * numsteps)
{
	if (s == NULL) return 1;
	for (int i = 0; i < GEVB_MAX_NCDM; i++)
	{
		if (maxvel) maxvel[i] = s->maxvel[2 + i];
		if (numsteps) numsteps[i] = s->numsteps_ncdm[i];
	}
	return 0;
}

extern "C" int gevb_sim_destroy(gevb_sim * s) { delete s; return 0; }

// the host-side Friedmann background by itself (no device needed): out = {Hconf(a), bg_ncdm(a), particleHorizon(a),
// a after rungekutta4bg(a, dtau)} -- background.hpp:103,137,167,200
extern "C" int gevb_background_eval(const double * c, int num_ncdm, const double * m_ncdm, const double * T_ncdm, const double * Omega_ncdm,
                                    double a, double fourpiG, double dtau, double * out4)
{
	if (c == NULL || out4 == NULL || num_ncdm < 0 || num_ncdm > GEVB_MAX_NCDM) return 1;
	cosmology co;
	std::memset(&co, 0, sizeof(co));
	co.Omega_cdm = c[0]; co.Omega_b = c[1]; co.Omega_m = c[2]; co.Omega_Lambda = c[3]; co.Omega_fld = c[4]; co.w0_fld = c[5]; co.wa_fld = c[6];
	co.Omega_g = c[7]; co.Omega_ur = c[8]; co.Omega_rad = c[9]; co.h = c[10]; co.num_ncdm = num_ncdm;
	for (int i = 0; i < num_ncdm; i++) { co.m_ncdm[i] = m_ncdm[i]; co.T_ncdm[i] = T_ncdm[i]; co.Omega_ncdm[i] = Omega_ncdm[i]; }
	out4[0] = Hconf(a, fourpiG, co);
	out4[1] = bg_ncdm(a, co);
	out4[2] = particleHorizon(a, fourpiG, co);
	double a2 = a;
	rungekutta4bg(a2, fourpiG, co, dtau);
	out4[3] = a2;
	return 0;
}

extern "C" int gevb_sim_set_particles(gevb_sim * s, int species, int64_t n, const int64_t * id, const double * pos, const double * vel, double mass)
{
	if (s == NULL || species < 0 || species >= 2 + GEVB_MAX_NCDM) return 1;
	part_simple_info info;
	info.mass = mass; info.relativistic = species >= 2; std::strcpy(info.type_name, "part_simple");
	Particles_gevolution & p = species == 0 ? s->pcls_cdm : (species == 1 ? s->pcls_b : s->pcls_ncdm[species - 2]);
	GEVB_C_BOUNDARY(p.initialize(info, &s->lat);)
	if (species == 1) s->baryon_flag = 1;
	return gevb_pcls_add(p.handle(), n, id, pos, vel);
}

extern "C" gevb_field * gevb_sim_field(gevb_sim * s, int which)
{
	switch (which)
	{
		case 0: return s->phi.handle(); case 1: return s->chi.handle(); case 2: return s->Bi.handle();
		case 3: return s->source.handle(); case 4: return s->Sij.handle();
		case 10: return s->scalarFT.handle(); case 11: return s->BiFT.handle(); case 12: return s->SijFT.handle();
	}
	return NULL;
}

extern "C" gevb_pcls * gevb_sim_pcls(gevb_sim * s, int species)
{
	if (s == NULL || species < 0 || species >= 2 + GEVB_MAX_NCDM) return NULL;
	return species == 0 ? s->pcls_cdm.handle() : (species == 1 ? s->pcls_b.handle() : s->pcls_ncdm[species - 2].handle());
}

extern "C" int gevb_sim_set_field(gevb_sim * s, int which, const double * host)
{
	gevb_field * f = gevb_sim_field(s, which);
	if (f == NULL) return 1;
	int r = gevb_field_upload(f, host);
	if (r == 0 && which < 10) r = gevb_field_updateHalo(f);
	return r;
}

extern "C" int gevb_sim_get_field(gevb_sim * s, int which, double * host)
{
	gevb_field * f = gevb_sim_field(s, which);
	return f ? gevb_field_download(f, host) : 1;
}

extern "C" int gevb_sim_get_state(gevb_sim * s, double * o)
{
	o[0] = s->a; o[1] = s->tau; o[2] = s->dtau; o[3] = s->dtau_old; o[4] = s->cycle;
	o[5] = s->maxvel[0]; o[6] = s->maxvel[1]; o[7] = s->T00hom; o[8] = s->fourpiG;
	return 0;
}

extern "C" int gevb_sim_set_state(gevb_sim * s, const double * in)
{
	s->a = in[0]; s->tau = in[1]; s->dtau = in[2]; s->dtau_old = in[3]; s->cycle = (int) in[4];
	s->maxvel[0] = in[5]; s->maxvel[1] = in[6];
	return 0;
}

// maxvel of the ncdm species as the IC generator returns it (main.cpp:330-340 feeds the first cycle's sub-stepping)
extern "C" int gevb_sim_set_ncdm_maxvel(gevb_sim * s, const double * maxvel)
{
	if (s == NULL || maxvel == NULL) return 1;
	for (int i = 0; i < GEVB_MAX_NCDM; i++) s->maxvel[2 + i] = maxvel[i];
	return 0;
}

extern "C" int gevb_sim_set_fused(gevb_sim * s, int fused)
{
	if (s == NULL) return 1;
	s->fused = fused;
	// fused mode treats scalarFT as scratch (its backward transforms may clobber it); unfused keeps LATfield2's semantics
	GEVB_C_BOUNDARY(s->plan_phi.preserveInput(!fused); s->plan_chi.preserveInput(!fused);)
	return 0;
}

static int sim_solve(gevb_sim * s);
static int sim_update(gevb_sim * s);
static int sim_step(gevb_sim * s) { int r = sim_solve(s); return r ? r : sim_update(s); }

// ---- hibernation / restart (hibernation.hpp:512-611 writes, ic_read.hpp:290-330 reads) ---------------------------
// The state set and the arithmetic are the reference's: the particles of every species, phi, chi, and the vector
// potential in REAL space divided by a^2 N (hibernation.hpp:533-538), plus the scalars a, tau, dtau, dtau_old, cycle,
// maxvel[] that it keeps in the restart settings file (writeRestartSettings).  At restart the stored B is multiplied
// by a^2 / N^2 and BiFT is rebuilt from it by a forward transform (ic_read.hpp:305-319).  The reference stores fields
// and particles as HDF5 (absent here): the fields go to flat binary files <filebase>_phi.bin, _chi.bin, _B.bin
// (gevb_field_save_raw), the particles and scalars of a rank to <filebase>.<rank>.gevb.
namespace {
const char HIB_MAGIC[8] = {'G', 'E', 'V', 'B', 'H', 'I', 'B', '2'};
struct HibHeader
{
	char magic[8];
	int32_t ngrid, nranks, rank, z0, nzl, nky, gr_flag, vector_flag, baryon_flag, num_ncdm, cycle, reserved;
	double a, tau, dtau, dtau_old, T00hom;
	double maxvel[2 + GEVB_MAX_NCDM];
	int64_t npart[2 + GEVB_MAX_NCDM];
	double mass[2 + GEVB_MAX_NCDM];
};
bool put(FILE * f, const void * p, size_t bytes) { return bytes == 0 || std::fwrite(p, 1, bytes, f) == bytes; }
bool get(FILE * f, void * p, size_t bytes) { return bytes == 0 || std::fread(p, 1, bytes, f) == bytes; }
Particles_gevolution * species_of(gevb_sim * s, int sp) { return sp == 0 ? &s->pcls_cdm : (sp == 1 ? &s->pcls_b : &s->pcls_ncdm[sp - 2]); }
struct FileCloser { FILE * f; ~FileCloser() { if (f) std::fclose(f); } };
}

static int sim_hibernate(gevb_sim * s, const char * filebase)
{
	gevb_ctx * ctx = s->lat.ctx();
	int rank = 0, nranks = 1, n, z0, nzl, ky0, nky;
	gevb_ctx_ranks(ctx, &rank, &nranks);
	gevb_ctx_geometry(ctx, &n, &z0, &nzl, &ky0, &nky);
	HibHeader h;
	std::memset(&h, 0, sizeof(h));
	std::memcpy(h.magic, HIB_MAGIC, 8);
	h.ngrid = n; h.nranks = nranks; h.rank = rank; h.z0 = z0; h.nzl = nzl; h.nky = nky;
	h.gr_flag = s->gr_flag; h.vector_flag = s->vector_flag; h.baryon_flag = s->baryon_flag; h.num_ncdm = s->cosmo.num_ncdm; h.cycle = s->cycle;
	h.a = s->a; h.tau = s->tau; h.dtau = s->dtau; h.dtau_old = s->dtau_old; h.T00hom = s->T00hom;
	for (int i = 0; i < 2 + GEVB_MAX_NCDM; i++)
	{
		h.maxvel[i] = s->maxvel[i];
		Particles_gevolution * p = species_of(s, i);
		h.npart[i] = p->initialized() ? p->numParticlesLocal() : -1;
		h.mass[i] = p->initialized() ? gevb_pcls_mass(p->handle()) : 0.;
	}
	char name[1024];
	bool ok = true;
	{
		std::snprintf(name, sizeof(name), "%s.%d.gevb", filebase, rank);
		FileCloser fc = {std::fopen(name, "wb")};
		ok = fc.f != NULL && put(fc.f, &h, sizeof(h));
		for (int i = 0; i < 2 + GEVB_MAX_NCDM && ok; i++)
		{
			if (h.npart[i] <= 0) continue;
			const size_t np = (size_t) h.npart[i];
			std::vector<int64_t> id(np);
			std::vector<double> pos(3 * np), vel(3 * np);
			ok = gevb_pcls_download(species_of(s, i)->handle(), id.data(), pos.data(), vel.data()) == 0
				&& put(fc.f, id.data(), np * 8) && put(fc.f, pos.data(), np * 24) && put(fc.f, vel.data(), np * 24);
		}
		if (fc.f) { ok = std::fclose(fc.f) == 0 && ok; fc.f = NULL; }
	}
	double bad = ok ? 0. : 1.;
	if (gevb_parallel_sum(ctx, &bad, 1) != 0 || bad != 0.) return 1;              // every rank enters the collective field writes below, or none
	// hibernation.hpp:533-538: the stored vector potential is Bi / (a^2 N).  As in the reference it is not scaled back: Bi is
	// derived state that the next metric solve rebuilds from BiFT (main.cpp:593) before anything reads it.
	if (s->vector_flag == VECTOR_PARABOLIC)
	{
		if (s->gr_flag == 0) s->plan_Bi.execute(FFT_BACKWARD);                    // main.cpp:837,850
		check(gevb_field_scale(s->Bi.handle(), 1. / (s->a * s->a * s->numpts)), "hibernate");
		std::snprintf(name, sizeof(name), "%s_B.bin", filebase);
		if (gevb_field_save_raw(s->Bi.handle(), name) != 0) return 1;             // hibernation.hpp:597
	}
	if (s->gr_flag > 0)
	{
		std::snprintf(name, sizeof(name), "%s_phi.bin", filebase);
		if (gevb_field_save_raw(s->phi.handle(), name) != 0) return 1;            // hibernation.hpp:591
		std::snprintf(name, sizeof(name), "%s_chi.bin", filebase);
		if (gevb_field_save_raw(s->chi.handle(), name) != 0) return 1;            // hibernation.hpp:592
	}
	else
	{
		// Newtonian runs recompute phi from the particles every cycle; it is written all the same so that a restart begins from the state left
		std::snprintf(name, sizeof(name), "%s_phi.bin", filebase);
		if (gevb_field_save_raw(s->phi.handle(), name) != 0) return 1;
	}
	return 0;
}

extern "C" int gevb_sim_hibernate(gevb_sim * s, const char * filebase)
{
	if (s == NULL || filebase == NULL) return 1;
	GEVB_C_BOUNDARY(return sim_hibernate(s, filebase);)
}

static int sim_restore(gevb_sim * s, const char * filebase);
extern "C" int gevb_sim_restore(gevb_sim * s, const char * filebase)
{
	if (s == NULL || filebase == NULL) return 1;
	GEVB_C_BOUNDARY(return sim_restore(s, filebase);)         // allocation failure and the like never cross the C boundary
}

static int sim_restore(gevb_sim * s, const char * filebase)
{
	gevb_ctx * ctx = s->lat.ctx();
	int rank = 0, nranks = 1, n, z0, nzl, ky0, nky;
	gevb_ctx_ranks(ctx, &rank, &nranks);
	gevb_ctx_geometry(ctx, &n, &z0, &nzl, &ky0, &nky);
	// 1. every rank reads and validates its whole particle file on the host; the ranks agree on the outcome before any
	//    device work or collective starts (a rank with a truncated file must not leave the others inside an exchange)
	HibHeader h;
	std::vector<int64_t> id[2 + GEVB_MAX_NCDM];
	std::vector<double> pos[2 + GEVB_MAX_NCDM], vel[2 + GEVB_MAX_NCDM];
	bool ok = true;
	try
	{
		char name[1024];
		std::snprintf(name, sizeof(name), "%s.%d.gevb", filebase, rank);
		FileCloser fc = {std::fopen(name, "rb")};
		ok = fc.f != NULL && get(fc.f, &h, sizeof(h)) && std::memcmp(h.magic, HIB_MAGIC, 8) == 0;
		if (ok)
		{
			// the particle counts must fit the file (a truncated or foreign file must not drive the allocations below)
			const long here = std::ftell(fc.f);
			std::fseek(fc.f, 0, SEEK_END);
			const long size = std::ftell(fc.f);
			std::fseek(fc.f, here, SEEK_SET);
			int64_t total = 0;
			for (int i = 0; i < 2 + GEVB_MAX_NCDM; i++) if (h.npart[i] > 0) total += h.npart[i];
			ok = here > 0 && size >= here && total >= 0 && total <= (int64_t) ((size - here) / 56);
		}
		// a restart must use the decomposition the state was written with (the reference has the same restriction per file set)
		ok = ok && h.ngrid == n && h.nranks == nranks && h.rank == rank && h.z0 == z0 && h.nzl == nzl && h.gr_flag == s->gr_flag && h.vector_flag == s->vector_flag
			&& h.num_ncdm == s->cosmo.num_ncdm;
		for (int i = 0; i < 2 + GEVB_MAX_NCDM && ok; i++)
		{
			if (h.npart[i] <= 0) continue;
			const size_t np = (size_t) h.npart[i];
			id[i].resize(np); pos[i].resize(3 * np); vel[i].resize(3 * np);
			ok = get(fc.f, id[i].data(), np * 8) && get(fc.f, pos[i].data(), np * 24) && get(fc.f, vel[i].data(), np * 24);
		}
	}
	catch (...) { ok = false; }
	double bad = ok ? 0. : 1.;
	if (gevb_parallel_sum(ctx, &bad, 1) != 0 || bad != 0.) return 1;
	// 2. fields (each load validates on the host and agrees across ranks before it uploads)
	char name[1024];
	std::snprintf(name, sizeof(name), "%s_phi.bin", filebase);
	if (gevb_field_load_raw(s->phi.handle(), name) != 0) return 1;                // ic_read.hpp (metricfile[0]) + updateHalo :288
	if (s->gr_flag > 0)
	{
		std::snprintf(name, sizeof(name), "%s_chi.bin", filebase);
		if (gevb_field_load_raw(s->chi.handle(), name) != 0) return 1;            // ic_read.hpp:323-327
	}
	s->a = h.a; s->tau = h.tau; s->dtau = h.dtau; s->dtau_old = h.dtau_old; s->cycle = h.cycle; s->T00hom = h.T00hom;
	for (int i = 0; i < 2 + GEVB_MAX_NCDM; i++) s->maxvel[i] = h.maxvel[i];
	if (s->vector_flag == VECTOR_PARABOLIC)
	{
		std::snprintf(name, sizeof(name), "%s_B.bin", filebase);
		if (gevb_field_load_raw(s->Bi.handle(), name) != 0) return 1;             // ic_read.hpp:309-310
		check(gevb_field_scale(s->Bi.handle(), s->a * s->a / ((double) s->numpts * (double) s->numpts)), "restore");   // :312-317
		s->plan_Bi.execute(FFT_FORWARD);                                          // :318: BiFT is rebuilt from the stored real-space field
	}
	// Bi in real space is derived state (main.cpp:593-598)
	if (s->gr_flag > 0) { s->plan_Bi.execute(FFT_BACKWARD); s->Bi.updateHalo(); }
	// 3. particles
	for (int i = 0; i < 2 + GEVB_MAX_NCDM; i++)
	{
		if (h.npart[i] < 0) continue;
		if (gevb_sim_set_particles(s, i, h.npart[i], id[i].data(), pos[i].data(), vel[i].data(), h.mass[i]) != 0) return 1;
	}
	return 0;
}

// writeSnapshots' field dumps (output.hpp:98-300) as flat binary files <prefix>_<T00|B|phi|chi|hij>.bin
static int write_field_snapshot(gevb_sim * s, const char * prefix, int mask)
{
	const double a = s->a;
	char name[1024];
	auto file = [&](const char * tag) { std::snprintf(name, sizeof(name), "%s_%s.bin", prefix, tag); return name; };
	if (mask & 16)                                                                           // MASK_T00, output.hpp:155-193
	{
		projection_init(&s->source);
		for (int sp = 0; sp < 2 + s->cosmo.num_ncdm; sp++)
		{
			Particles_gevolution * p = species_of(s, sp);
			if (!p->initialized()) continue;
			if (s->gr_flag > 0) projection_T00_project(p, &s->source, a, &s->phi);           // :160-167
			else scalarProjectionCIC_project(p, &s->source);                                 // :171-178
		}
		projection_T00_comm(&s->source);                                                     // :181
		if (gevb_field_save_raw(s->source.handle(), file("T00")) != 0) return 1;
	}
	if (mask & 8)                                                                            // MASK_B, output.hpp:206-237
	{
		if (s->gr_flag == 0) s->plan_Bi.execute(FFT_BACKWARD);                               // :210
		check(gevb_field_scale(s->Bi.handle(), 1. / (a * a * s->numpts)), "writeSnapshots"); // :212-217
		s->Bi.updateHalo();                                                                  // :218
		if (gevb_field_save_raw(s->Bi.handle(), file("B")) != 0) return 1;                   // :229
		if (s->gr_flag > 0) { s->plan_Bi.execute(FFT_BACKWARD); s->Bi.updateHalo(); }        // :232-236: restored from BiFT
	}
	if ((mask & 1) && gevb_field_save_raw(s->phi.handle(), file("phi")) != 0) return 1;      // MASK_PHI, :239-247
	if ((mask & 2) && gevb_field_save_raw(s->chi.handle(), file("chi")) != 0) return 1;      // MASK_CHI, :249-257
	if (mask & 128)                                                                          // MASK_HIJ, :259-277 (done_hij == 0)
	{
		projectFTtensor(s->SijFT, s->SijFT);
		s->plan_Sij.execute(FFT_BACKWARD);
		s->Sij.updateHalo();
		if (gevb_field_save_raw(s->Sij.handle(), file("hij")) != 0) return 1;
	}
	return 0;
}

extern "C" int gevb_sim_write_field_snapshot(gevb_sim * s, const char * prefix, int mask)
{
	if (s == NULL || prefix == NULL) return 1;
	GEVB_C_BOUNDARY(return write_field_snapshot(s, prefix, mask);)
}

// the phi / chi / hij / B part of writeSpectra (output.hpp:1945-1981,2151-2155; call main.cpp:639-679): forward
// transforms, TT projection for hij, binning on the device, one text file per spectrum named <prefix><pkcount>_<name>.dat
static int write_spectra(gevb_sim * s, const char * prefix, int pkcount, int numbins, int mask, double z_target)
{
	const double a = s->a, fourpiG = s->fourpiG;
	const double numpts3d = (double) s->numpts * (double) s->numpts * (double) s->numpts;
	std::vector<double> kbin(numbins), power(numbins), kscatter(numbins), pscatter(numbins);
	std::vector<int> occupation(numbins);
	char filename[1024];
	int rank = 0;
	gevb_ctx_ranks(s->lat.ctx(), &rank, NULL);
	auto emit = [&](Field<Cplx> & fld, const char * tag, double rescalep, const char * description)
	{
		extractPowerSpectrum(fld, kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, false, 1 /* KTYPE_LINEAR */);
		std::snprintf(filename, sizeof(filename), "%s%03d_%s.dat", prefix, pkcount, tag);
		if (rank == 0)                                                                       // parallel.isRoot(), tools.hpp:270
			check(gevb_writePowerSpectrum(kbin.data(), power.data(), kscatter.data(), pscatter.data(), occupation.data(), numbins, s->boxsize, rescalep, filename, description, a, z_target), "writePowerSpectrum");
	};
	if (mask & 1)                                                                            // MASK_PHI, output.hpp:1945-1951
	{
		s->plan_phi.execute(FFT_FORWARD);
		emit(s->scalarFT, "phi", numpts3d * numpts3d * 2. * M_PI * M_PI, "power spectrum of phi");
	}
	if (mask & 2)                                                                            // MASK_CHI, :1953-1959
	{
		s->plan_chi.execute(FFT_FORWARD);
		emit(s->scalarFT, "chi", numpts3d * numpts3d * 2. * M_PI * M_PI, "power spectrum of chi");
	}
	if (mask & 128)                                                                          // MASK_HIJ, :1961-1981
	{
		projection_init(&s->Sij);
		projection_Tij_project(&s->pcls_cdm, &s->Sij, a, &s->phi);
		if (s->baryon_flag) projection_Tij_project(&s->pcls_b, &s->Sij, a, &s->phi);
		for (int i = 0; i < s->cosmo.num_ncdm; i++)
			if (s->pcls_ncdm[i].initialized()) projection_Tij_project(s->pcls_ncdm + i, &s->Sij, a, &s->phi);
		projection_Tij_comm(&s->Sij);
		prepareFTsource<Real>(s->phi, s->Sij, s->Sij, 2. * fourpiG / (double) s->numpts / (double) s->numpts / a);
		s->plan_Sij.execute(FFT_FORWARD);
		projectFTtensor(s->SijFT, s->SijFT);
		emit(s->SijFT, "hij", 2. * M_PI * M_PI, "power spectrum of hij");
	}
	if (mask & 8)                                                                            // MASK_B, :2151-2155
		emit(s->BiFT, "B", a * a * a * a * s->numpts * s->numpts * 2. * M_PI * M_PI, "power spectrum of B");
	return 0;
}

extern "C" int gevb_sim_write_spectra(gevb_sim * s, const char * prefix, int pkcount, int numbins, int mask, double z_target)
{
	if (s == NULL || prefix == NULL || numbins < 1) return 1;
	GEVB_C_BOUNDARY(return write_spectra(s, prefix, pkcount, numbins, mask, z_target);)
}

// snapshot of one species in Gadget-2 format (writeSnapshots, output.hpp:62-470 -> saveGadget2): the header is filled
// as output.hpp:360-402 does (mass in 1e10 M_sun/h from the critical density, box in kpc/h)
static int sim_save_gadget2(gevb_sim * s, int species, const char * filename, int tracer_factor, double dtau_pos, double dtau_vel, double time, double redshift);

extern "C" int gevb_sim_save_gadget2(gevb_sim * s, int species, const char * filename, int tracer_factor, double dtau_pos, double dtau_vel)
{
	if (s == NULL) return 1;
	return sim_save_gadget2(s, species, filename, tracer_factor, dtau_pos, dtau_vel, s->a, (1. / s->a) - 1.);      // stamped with the current scale factor (output.hpp:390-391)
}

// `time`: the scale factor the file is stamped with and the velocities are scaled by (the target of the half-step corrections)
static int sim_save_gadget2(gevb_sim * s, int species, const char * filename, int tracer_factor, double dtau_pos, double dtau_vel, double time, double redshift)
{
	if (s == NULL || filename == NULL) return 1;
	gevb_pcls * p = gevb_sim_pcls(s, species);
	if (p == NULL) return 1;
	struct { uint32_t npart[6]; double mass[6]; double time, redshift; int32_t flag_sfr, flag_feedback; uint32_t npartTotal[6]; int32_t flag_cooling, num_files;
	         double BoxSize, Omega0, OmegaLambda, HubbleParam; int32_t flag_age, flag_metals; uint32_t npartTotalHW[6]; char fill[64]; } hdr;   // metadata.hpp:152-171
	static_assert(sizeof(hdr) == 256, "gadget2 header must be 256 bytes");
	std::memset(&hdr, 0, sizeof(hdr));
	hdr.num_files = 1;
	hdr.Omega0 = s->cosmo.Omega_m; hdr.OmegaLambda = s->cosmo.Omega_Lambda; hdr.HubbleParam = s->cosmo.h;
	hdr.BoxSize = s->boxsize / 0.001;                                                        // GADGET_LENGTH_CONVERSION, output.hpp:369
	hdr.time = time; hdr.redshift = redshift;
	int64_t n_local = 0;
	gevb_pcls_count(p, &n_local);
	double ntot = (double) n_local;
	if (gevb_parallel_sum(s->lat.ctx(), &ntot, 1) != 0) return 1;
	const int64_t nsel = ((int64_t) ntot + tracer_factor - 1) / tracer_factor;               // output.hpp:396 (saveGadget2 overwrites npart[1] with the count it finds)
	hdr.npart[1] = (uint32_t) (nsel % (1ll << 32)); hdr.npartTotal[1] = hdr.npart[1]; hdr.npartTotalHW[1] = (uint32_t) (nsel / (1ll << 32));
	// C_RHO_CRIT = 2.77459457e11 (M_sun/h) / (Mpc/h)^3, metadata.hpp:99; species mass fraction times the box mass
	hdr.mass[1] = (double) tracer_factor * 2.77459457e11 * gevb_pcls_mass(p) * s->boxsize * s->boxsize * s->boxsize / 1.0e10;   // output.hpp:400-402,417,433
	return gevb_pcls_saveGadget2(p, filename, &hdr, tracer_factor, dtau_pos, dtau_vel, s->phi.handle());
}

// the main loop with its snapshot and power-spectrum outputs (main.cpp:372-879; lightcones and HDF5 dumps are not built):
// cycles run until every requested output is written (main.cpp:685-693) or max_cycles is reached.  Outputs sit between the
// metric solve and the particle update, as in the reference:
//   snapshot when 1/a < z_snapshot + 1 (:617-633): Gadget-2 file <snap_prefix><count %03d>_cdm (and _b), stamped with the
//     target redshift and drifted / kicked to it (EXACT_OUTPUT_REDSHIFTS, output.hpp:384-388,409);
//   spectra when 1/a < z_pk + 1 (:641-659), and once more one cycle before the target is passed (:661-679) so that
//     writePowerSpectrum can interpolate to the exact redshift.
extern "C" int gevb_sim_run(gevb_sim * s, const double * z_pk, int num_pk, int pk_mask, int numbins, const char * pk_prefix,
                            const double * z_snapshot, int num_snapshot, int tracer_factor, const char * snap_prefix, int max_cycles, int * counts3)
{
	if (s == NULL || !s->pcls_cdm.initialized() || (num_pk > 0 && (z_pk == NULL || pk_prefix == NULL)) || (num_snapshot > 0 && (z_snapshot == NULL || snap_prefix == NULL))) return 1;
	int pkcount = 0, snapcount = 0, cycles = 0;
	try
	{
		while (cycles < max_cycles)
		{
			if (sim_solve(s) != 0) return 1;
			const double a = s->a;
			if (snapcount < num_snapshot && 1. / a < z_snapshot[snapcount] + 1.)                                     // main.cpp:617
			{
				const double time = 1. / (z_snapshot[snapcount] + 1.);                                               // output.hpp:386
				const double dtau_pos = (time - a) / a / Hconf(a, s->fourpiG, s->cosmo);                             // output.hpp:388
				char name[1024];
				for (int sp = 0; sp < 2 + s->cosmo.num_ncdm; sp++)
				{
					Particles_gevolution & p = sp == 0 ? s->pcls_cdm : (sp == 1 ? s->pcls_b : s->pcls_ncdm[sp - 2]);
					if (!p.initialized()) continue;
					if (sp == 0) std::snprintf(name, sizeof(name), "%s%03d_cdm", snap_prefix, snapcount);
					else if (sp == 1) std::snprintf(name, sizeof(name), "%s%03d_b", snap_prefix, snapcount);
					else std::snprintf(name, sizeof(name), "%s%03d_ncdm%d", snap_prefix, snapcount, sp - 2);
					if (sim_save_gadget2(s, sp, name, tracer_factor, dtau_pos, dtau_pos + 0.5 * s->dtau_old, time, z_snapshot[snapcount]) != 0) return 1;   // output.hpp:386-387,409,423,439
				}
				snapcount++;
			}
			if (pkcount < num_pk && 1. / a < z_pk[pkcount] + 1.)                                                     // main.cpp:641
			{
				if (write_spectra(s, pk_prefix, pkcount, numbins, pk_mask, z_pk[pkcount]) != 0) return 1;
				pkcount++;
			}
			double tmp = a;                                                                                          // main.cpp:661-664
			rungekutta4bg(tmp, s->fourpiG, s->cosmo, 0.5 * s->dtau);
			rungekutta4bg(tmp, s->fourpiG, s->cosmo, 0.5 * s->dtau);
			if (pkcount < num_pk && 1. / tmp < z_pk[pkcount] + 1.)                                                   // main.cpp:666
				if (write_spectra(s, pk_prefix, pkcount, numbins, pk_mask, z_pk[pkcount]) != 0) return 1;
			if (pkcount >= num_pk && snapcount >= num_snapshot) break;                                               // main.cpp:685-693: simulation complete
			if (sim_update(s) != 0) return 1;
			cycles++;
		}
	}
	catch (const gevb_error &) { return 1; }
	catch (...) { return 1; }                                                                                    // std::bad_alloc of the output buffers and the like
	if (counts3) { counts3[0] = cycles; counts3[1] = pkcount; counts3[2] = snapcount; }
	return 0;
}

// one cycle of the main loop (main.cpp:372-879 without outputs); errors come back as a status
extern "C" int gevb_sim_step(gevb_sim * s)
{
	if (s == NULL || !s->pcls_cdm.initialized()) return 1;
	GEVB_C_BOUNDARY(return sim_step(s);)
}

// first half of a cycle: stress-energy projections and the metric solve (main.cpp:378-599); the outputs of the reference sit
// between the two halves (main.cpp:605-679)
static int sim_solve(gevb_sim * s)
{
	const double dx = s->dx, fourpiG = s->fourpiG;
	cosmology & cosmo = s->cosmo;
	double & a = s->a; double & dtau_old = s->dtau_old;
	Field<Real> & phi = s->phi, & chi = s->chi, & source = s->source, & Sij = s->Sij, & Bi = s->Bi;
	Field<Cplx> & scalarFT = s->scalarFT, & SijFT = s->SijFT, & BiFT = s->BiFT;
	Particles_gevolution & pcls_cdm = s->pcls_cdm, & pcls_b = s->pcls_b;
	Particles_gevolution * pcls_ncdm = s->pcls_ncdm;
	bool ncdm_T00[GEVB_MAX_NCDM], ncdm_Tij[GEVB_MAX_NCDM];                                    // which ncdm species deposit this cycle
	for (int i = 0; i < GEVB_MAX_NCDM; i++)
	{
		const bool have = i < cosmo.num_ncdm && pcls_ncdm[i].initialized();                   // sim.numpcl[1+sim.baryon_flag+i] > 0
		ncdm_T00[i] = have && a >= 1. / (s->z_switch_deltancdm[i] + 1.);                      // :390
		ncdm_Tij[i] = have && a >= 1. / (s->z_switch_linearchi + 1.);                         // :442-447
	}
	const bool fuse = s->fused && s->gr_flag > 0;

	// construct stress-energy tensor (main.cpp:378-450)
	projection_init(&source);
	projection_init(&Sij);
	if (fuse)
	{
		// one pass over the particles deposits T00 and Tij (same sums as :385 and :439)
		projection_T00_Tij_project(&pcls_cdm, &source, &Sij, a, &phi);
		if (s->baryon_flag) projection_T00_Tij_project(&pcls_b, &source, &Sij, a, &phi);
		for (int i = 0; i < cosmo.num_ncdm; i++)
		{
			if (ncdm_T00[i] && ncdm_Tij[i]) projection_T00_Tij_project(pcls_ncdm + i, &source, &Sij, a, &phi);
			else if (ncdm_T00[i]) projection_T00_project(pcls_ncdm + i, &source, a, &phi);
			else if (ncdm_Tij[i]) projection_Tij_project(pcls_ncdm + i, &Sij, a, &phi);
		}
	}
	else if (s->gr_flag > 0)
	{
		projection_T00_project(&pcls_cdm, &source, a, &phi);                                  // :385
		if (s->baryon_flag) projection_T00_project(&pcls_b, &source, a, &phi);                // :387
		for (int i = 0; i < cosmo.num_ncdm; i++)
			if (ncdm_T00[i]) projection_T00_project(pcls_ncdm + i, &source, a, &phi);         // :391
	}
	else
	{
		scalarProjectionCIC_project(&pcls_cdm, &source);                                      // :402
		if (s->baryon_flag) scalarProjectionCIC_project(&pcls_b, &source);                    // :404
		for (int i = 0; i < cosmo.num_ncdm; i++)
			if (ncdm_T00[i]) scalarProjectionCIC_project(pcls_ncdm + i, &source);             // :408
	}
	if (s->gr_flag > 0)
		for (int i = 0; i < cosmo.num_ncdm; i++)
			if (!ncdm_T00[i])                                                                 // :392-397 (radiation_flag == 0): homogeneous stand-in
				check(gevb_field_add_constant(source.handle(), 0, bg_ncdm(a, cosmo, i)), "bg_ncdm");
	projection_T00_comm(&source);                                                             // :411

	if (s->vector_flag == VECTOR_ELLIPTIC)
	{
		projection_init(&Bi);                                                                 // :426
		projection_T0i_project(&pcls_cdm, &Bi, &phi);                                         // :427
		if (s->baryon_flag) projection_T0i_project(&pcls_b, &Bi, &phi);                       // :429
		for (int i = 0; i < cosmo.num_ncdm; i++)
			if (pcls_ncdm[i].initialized() && a >= 1. / (s->z_switch_Bncdm[i] + 1.)) projection_T0i_project(pcls_ncdm + i, &Bi, &phi);   // :432-433
		projection_T0i_comm(&Bi);                                                             // :435
	}

	if (!fuse)
	{
		projection_Tij_project(&pcls_cdm, &Sij, a, &phi);                                     // :439
		if (s->baryon_flag) projection_Tij_project(&pcls_b, &Sij, a, &phi);                   // :441
		for (int i = 0; i < cosmo.num_ncdm; i++)
			if (ncdm_Tij[i]) projection_Tij_project(pcls_ncdm + i, &Sij, a, &phi);            // :442-448
	}
	projection_Tij_comm(&Sij);                                                                // :450

	if (s->gr_flag > 0)
	{
		double T00hom = 0.;
		const bool fuse_sum = fuse && dtau_old > 0.;                                          // the sum rides on prepareFTsource's pass
		if (!fuse_sum) check(gevb_field_sum(source.handle(), 0, &T00hom), "T00hom");          // :459-462 (sum + parallel.sum)

		if (dtau_old > 0.)
		{
			if (fuse)                                                                         // :472 + :477 in one pass where the own x-pass applies
				prepareFTsource_execute(phi, chi, s->plan_source, cosmo.Omega_cdm + cosmo.Omega_b + bg_ncdm(a, cosmo), 3. * Hconf(a, fourpiG, cosmo) * dx * dx / dtau_old, fourpiG * dx * dx / a, 3. * Hconf(a, fourpiG, cosmo) * Hconf(a, fourpiG, cosmo) * dx * dx, fuse_sum ? &T00hom : NULL);
			else
			{
				prepareFTsource<Real>(phi, chi, source, cosmo.Omega_cdm + cosmo.Omega_b + bg_ncdm(a, cosmo), source, 3. * Hconf(a, fourpiG, cosmo) * dx * dx / dtau_old, fourpiG * dx * dx / a, 3. * Hconf(a, fourpiG, cosmo) * Hconf(a, fourpiG, cosmo) * dx * dx);   // :472
				s->plan_source.execute(FFT_FORWARD);                                          // :477
			}
			solveModifiedPoissonFT(scalarFT, scalarFT, 1. / (dx * dx), 3. * Hconf(a, fourpiG, cosmo) / dtau_old);   // :483
			s->plan_phi.execute(FFT_BACKWARD);                                                // :488
		}
		T00hom /= (double) ((long) s->numpts * (long) s->numpts * (long) s->numpts);          // :463
		s->T00hom = T00hom;
	}
	else
	{
		s->plan_source.execute(FFT_FORWARD);                                                  // :500
		solveModifiedPoissonFT(scalarFT, scalarFT, fourpiG / a);                              // :506
		s->plan_phi.execute(FFT_BACKWARD);                                                    // :511
	}

	phi.updateHalo();                                                                         // :518

	if (fuse) prepareFTsource_execute(phi, s->plan_Sij, 2. * fourpiG * dx * dx / a);          // :539 + :544
	else
	{
		prepareFTsource<Real>(phi, Sij, Sij, 2. * fourpiG * dx * dx / a);                     // :539
		s->plan_Sij.execute(FFT_FORWARD);                                                     // :544
	}
	// parabolic B: :558 and :586 both read SijFT and nothing in between writes it, so one pass over it serves both
	const bool fuse_chi_B = fuse && s->vector_flag != VECTOR_ELLIPTIC;
	if (fuse_chi_B) projectFTscalar_evolveFTvector(SijFT, scalarFT, BiFT, a * a * dtau_old);  // :558 + :586
	else projectFTscalar(SijFT, scalarFT);                                                    // :558
	s->plan_chi.execute(FFT_BACKWARD);                                                        // :563
	chi.updateHalo();                                                                         // :568

	if (s->vector_flag == VECTOR_ELLIPTIC)
	{
		s->plan_Bi.execute(FFT_FORWARD);                                                      // :575
		projectFTvector(BiFT, BiFT, fourpiG * dx * dx);                                       // :580
	}
	else if (!fuse_chi_B)
		evolveFTvector(SijFT, BiFT, a * a * dtau_old);                                        // :586

	if (s->gr_flag > 0)
	{
		s->plan_Bi.execute(FFT_BACKWARD);                                                     // :593
		Bi.updateHalo();                                                                      // :598
	}

	return 0;
}

// second half of a cycle: particle updates, background, next time step (main.cpp:696-875)
static int sim_update(gevb_sim * s)
{
	const double dx = s->dx, fourpiG = s->fourpiG;
	cosmology & cosmo = s->cosmo;
	double & a = s->a; double & dtau = s->dtau; double & dtau_old = s->dtau_old;
	Field<Real> & phi = s->phi, & chi = s->chi, & Bi = s->Bi;
	Particles_gevolution & pcls_cdm = s->pcls_cdm, & pcls_b = s->pcls_b;
	Particles_gevolution * pcls_ncdm = s->pcls_ncdm;
	Field<Real> * update_cdm_fields[3] = {&phi, &chi, &Bi};
	Field<Real> ** update_ncdm_fields = update_cdm_fields;                                    // main.cpp:255-261: the same three fields
	double f_params[5];
	const bool fuse = s->fused && s->gr_flag > 0;

	// number of step subdivisions for the ncdm particle updates (main.cpp:696-701)
	for (int i = 0; i < cosmo.num_ncdm; i++)
	{
		if (dtau * s->maxvel[2 + i] > dx * s->movelimit)
			s->numsteps_ncdm[i] = (int) ceil(dtau * s->maxvel[2 + i] / dx / s->movelimit);
		else s->numsteps_ncdm[i] = 1;
	}

	// non-cold DM particle update (main.cpp:729-765)
	for (int i = 0; i < cosmo.num_ncdm; i++)
	{
		if (!pcls_ncdm[i].initialized()) continue;
		double tmp = a;
		const int numsteps = s->numsteps_ncdm[i];
		for (int j = 0; j < numsteps; j++)
		{
			f_params[0] = tmp;
			f_params[1] = tmp * tmp * s->numpts;
			if (fuse)
			{
				double tmp_half = tmp;
				rungekutta4bg(tmp_half, fourpiG, cosmo, 0.5 * dtau / numsteps);               // :748
				double d_params[2] = {tmp_half, tmp_half * tmp_half * s->numpts};
				const int nf = (1. / a < s->z_relax + 1. ? 3 : 2);
				s->maxvel[2 + i] = pcls_ncdm[i].kickDrift(update_q, (dtau + dtau_old) / 2. / numsteps, nf, f_params, dtau / numsteps, nf, d_params, update_ncdm_fields);
				tmp = tmp_half;
			}
			else
			{
				if (s->gr_flag > 0)
					s->maxvel[2 + i] = pcls_ncdm[i].updateVel(update_q, (dtau + dtau_old) / 2. / numsteps, update_ncdm_fields, (1. / a < s->z_relax + 1. ? 3 : 2), f_params);   // :738
				else
					s->maxvel[2 + i] = pcls_ncdm[i].updateVel(update_q_Newton, (dtau + dtau_old) / 2. / numsteps, update_ncdm_fields, 1, f_params);   // :740 (no radiation / fluid)
				rungekutta4bg(tmp, fourpiG, cosmo, 0.5 * dtau / numsteps);                    // :748
				f_params[0] = tmp;
				f_params[1] = tmp * tmp * s->numpts;
				if (s->gr_flag > 0)
					pcls_ncdm[i].moveParticles(update_pos, dtau / numsteps, update_ncdm_fields, (1. / a < s->z_relax + 1. ? 3 : 2), f_params);   // :753
				else
					pcls_ncdm[i].moveParticles(update_pos_Newton, dtau / numsteps, NULL, 0, f_params);   // :755
			}
			rungekutta4bg(tmp, fourpiG, cosmo, 0.5 * dtau / numsteps);                        // :761
		}
	}

	// cdm and baryon particle update (main.cpp:771-807)
	f_params[0] = a;
	f_params[1] = a * a * s->numpts;
	if (fuse)
	{
		// kick uses {a, a^2 N}; the drift uses the scale factor after half a step (:792-795)
		double a_half = a;
		rungekutta4bg(a_half, fourpiG, cosmo, 0.5 * dtau);
		double d_params[2] = {a_half, a_half * a_half * s->numpts};
		const int nf_kick = (1. / a < s->z_relax + 1. ? 3 : 2), nf_drift = (1. / a_half < s->z_relax + 1. ? 3 : 0);
		s->maxvel[0] = pcls_cdm.kickDrift(update_q, (dtau + dtau_old) / 2., nf_kick, f_params, dtau, nf_drift, d_params, update_cdm_fields);
		if (s->baryon_flag) s->maxvel[1] = pcls_b.kickDrift(update_q, (dtau + dtau_old) / 2., nf_kick, f_params, dtau, nf_drift, d_params, update_cdm_fields);
		a = a_half;
	}
	else
	{
		if (s->gr_flag > 0)
		{
			s->maxvel[0] = pcls_cdm.updateVel(update_q, (dtau + dtau_old) / 2., update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 2), f_params);   // :775
			if (s->baryon_flag) s->maxvel[1] = pcls_b.updateVel(update_q, (dtau + dtau_old) / 2., update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 2), f_params);
		}
		else
		{
			s->maxvel[0] = pcls_cdm.updateVel(update_q_Newton, (dtau + dtau_old) / 2., update_cdm_fields, 1, f_params);   // :781
			if (s->baryon_flag) s->maxvel[1] = pcls_b.updateVel(update_q_Newton, (dtau + dtau_old) / 2., update_cdm_fields, 1, f_params);
		}

		rungekutta4bg(a, fourpiG, cosmo, 0.5 * dtau);                                         // :792

		f_params[0] = a;
		f_params[1] = a * a * s->numpts;
		if (s->gr_flag > 0)
		{
			pcls_cdm.moveParticles(update_pos, dtau, update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 0), f_params);   // :798
			if (s->baryon_flag) pcls_b.moveParticles(update_pos, dtau, update_cdm_fields, (1. / a < s->z_relax + 1. ? 3 : 0), f_params);
		}
		else
		{
			pcls_cdm.moveParticles(update_pos_Newton, dtau, NULL, 0, f_params);               // :804
			if (s->baryon_flag) pcls_b.moveParticles(update_pos_Newton, dtau, NULL, 0, f_params);
		}
	}

	rungekutta4bg(a, fourpiG, cosmo, 0.5 * dtau);                                             // :814

	s->lat.max(s->maxvel, 2 + cosmo.num_ncdm);                                                // :816 (numspecies entries; an absent baryon slot stays 0)
	if (s->gr_flag > 0)
		for (int i = 0; i < 2 + cosmo.num_ncdm; i++) s->maxvel[i] /= sqrt(s->maxvel[i] * s->maxvel[i] + 1.0);   // :818-822

	s->tau += dtau;                                                                           // :825
	dtau_old = dtau;                                                                          // :867
	if (s->Cf * dx < s->steplimit / Hconf(a, fourpiG, cosmo)) dtau = s->Cf * dx;              // :869-872
	else dtau = s->steplimit / Hconf(a, fourpiG, cosmo);
	s->cycle++;                                                                               // :874
	return 0;
}
