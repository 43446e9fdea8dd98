// Compile-only check of the drop-in boundary: the statements of the reference's time loop
// (main.cpp:372-879 and the IC calls ic_basic.hpp:1995,2064), written as the reference writes them -- same function
// names, argument order and types -- must compile against include/gevolution_b200.hpp.  Only the set-up lines differ
// (device / rank instead of parallel.initialize), as INTEGRATION.md describes.  tests/test_abi.py runs
// `g++ -fsyntax-only` on this file; nothing here is executed.
#include "gevolution_b200.hpp"
using namespace gevb200;

#define VECTOR_ELLIPTIC 1

struct metadata { int numpts, gr_flag, vector_flag, baryon_flag; double movelimit; };
struct icsettings { double z_relax; };
struct cosmology_t { double Omega_cdm, Omega_b; int num_ncdm; };

static double Hconf(double, double, const cosmology_t &) { return 1.; }
static double bg_ncdm(double, const cosmology_t &) { return 0.; }
static void rungekutta4bg(double &, double, const cosmology_t &, double) {}

void one_cycle(Lattice & lat, metadata & sim, icsettings & ic, cosmology_t & cosmo)
{
	// main.cpp:217-246
	Particles_gevolution pcls_cdm, pcls_b, pcls_ncdm[4];
	Field<Real> phi, source, chi, Sij, Bi;
	Field<Cplx> scalarFT, SijFT, BiFT;
	source.initialize(lat, 1); phi.initialize(lat, 1); chi.initialize(lat, 1); scalarFT.initialize(lat, 1);
	PlanFFT<Cplx> plan_source(&source, &scalarFT), plan_phi(&phi, &scalarFT), plan_chi(&chi, &scalarFT);
	Sij.initialize(lat, 3, 3, symmetric); SijFT.initialize(lat, 3, 3, symmetric);
	PlanFFT<Cplx> plan_Sij(&Sij, &SijFT);
	Bi.initialize(lat, 3); BiFT.initialize(lat, 3);
	PlanFFT<Cplx> plan_Bi(&Bi, &BiFT);
	Field<Real> * update_cdm_fields[3] = {&phi, &chi, &Bi};
	Field<Real> * ic_fields[2] = {&chi, &phi};
	double f_params[5], maxvel[6], a = 0.01, dtau = 0.1, dtau_old = 0.1, fourpiG = 1., dx = 1. / sim.numpts, T00hom = 0., max_displacement;
	int i = MAX;

	// ic_basic.hpp:1995,2064
	pcls_cdm.moveParticles(displace_pcls_ic_basic, 1., ic_fields, 2, NULL, &max_displacement, &i, 1);
	maxvel[0] = pcls_cdm.updateVel(initialize_q_ic_basic, 1., ic_fields, 2) / a;

	// main.cpp:378-450
	projection_init(&source);
	projection_T00_project(&pcls_cdm, &source, a, &phi);
	if (sim.baryon_flag) projection_T00_project(&pcls_b, &source, a, &phi);
	for (i = 0; i < cosmo.num_ncdm; i++) projection_T00_project(pcls_ncdm + i, &source, a, &phi);
	scalarProjectionCIC_project(&pcls_cdm, &source);
	projection_T00_comm(&source);
	projection_init(&Bi);
	projection_T0i_project(&pcls_cdm, &Bi, &phi);
	projection_T0i_comm(&Bi);
	projection_init(&Sij);
	projection_Tij_project(&pcls_cdm, &Sij, a, &phi);
	projection_Tij_comm(&Sij);

	// main.cpp:459-518
	gevb_field_sum(source.handle(), 0, &T00hom);          // the site loop + parallel.sum of :459-462
	prepareFTsource<Real>(phi, chi, source, cosmo.Omega_cdm + cosmo.Omega_b + bg_ncdm(a, cosmo), source, 3. * Hconf(a, fourpiG, cosmo) * dx * dx / dtau_old, fourpiG * dx * dx / a, 3. * Hconf(a, fourpiG, cosmo) * Hconf(a, fourpiG, cosmo) * dx * dx);
	plan_source.execute(FFT_FORWARD);
	solveModifiedPoissonFT(scalarFT, scalarFT, 1. / (dx * dx), 3. * Hconf(a, fourpiG, cosmo) / dtau_old);
	plan_phi.execute(FFT_BACKWARD);
	solveModifiedPoissonFT(scalarFT, scalarFT, fourpiG / a);
	phi.updateHalo();

	// main.cpp:539-598
	prepareFTsource<Real>(phi, Sij, Sij, 2. * fourpiG * dx * dx / a);
	plan_Sij.execute(FFT_FORWARD);
	projectFTscalar(SijFT, scalarFT);
	plan_chi.execute(FFT_BACKWARD);
	chi.updateHalo();
	if (sim.vector_flag == VECTOR_ELLIPTIC)
	{
		plan_Bi.execute(FFT_FORWARD);
		projectFTvector(BiFT, BiFT, fourpiG * dx * dx);
	}
	else
		evolveFTvector(SijFT, BiFT, a * a * dtau_old);
	plan_Bi.execute(FFT_BACKWARD);
	Bi.updateHalo();
	projectFTtensor(SijFT, SijFT);                        // output.hpp:1976

	// main.cpp:771-822
	f_params[0] = a;
	f_params[1] = a * a * sim.numpts;
	if (sim.gr_flag > 0)
		maxvel[0] = pcls_cdm.updateVel(update_q, (dtau + dtau_old) / 2., update_cdm_fields, (1. / a < ic.z_relax + 1. ? 3 : 2), f_params);
	else
		maxvel[0] = pcls_cdm.updateVel(update_q_Newton, (dtau + dtau_old) / 2., update_cdm_fields, 1, f_params);
	rungekutta4bg(a, fourpiG, cosmo, 0.5 * dtau);
	f_params[0] = a;
	f_params[1] = a * a * sim.numpts;
	if (sim.gr_flag > 0)
		pcls_cdm.moveParticles(update_pos, dtau, update_cdm_fields, (1. / a < ic.z_relax + 1. ? 3 : 0), f_params);
	else
		pcls_cdm.moveParticles(update_pos_Newton, dtau, NULL, 0, f_params);
	rungekutta4bg(a, fourpiG, cosmo, 0.5 * dtau);
	lat.max(maxvel, 1 + sim.baryon_flag);                 // parallel.max<double>(maxvel, numspecies), :816

	// tools.hpp:237 (writeSpectra)
	Real kbin[16], power[16], kscatter[16], pscatter[16];
	int occupation[16];
	extractPowerSpectrum(scalarFT, kbin, power, kscatter, pscatter, occupation, 16, false, 1);
}
