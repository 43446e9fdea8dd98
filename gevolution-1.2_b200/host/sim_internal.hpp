// sim_internal.hpp -- state of one simulation (main.cpp:217-297), shared by the host-side sources
// (sim.cpp: the time loop; settings.cpp: settings.ini; ic_basic.cpp: the basic IC generator)
#ifndef GEVB_HOST_SIM_INTERNAL_HPP
#define GEVB_HOST_SIM_INTERNAL_HPP
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#define GEVB_THROW_ON_ERROR
#include "../../include/gevolution_b200.hpp"
#include "background.hpp"

using namespace gevb200;

#define VECTOR_PARABOLIC 0      // metadata.hpp:92
#define VECTOR_ELLIPTIC 1       // metadata.hpp:93

struct gevb_sim
{
	Lattice lat;
	cosmology cosmo;
	int numpts, gr_flag, vector_flag, baryon_flag, fused;
	double boxsize, Cf, steplimit, z_in, z_relax;
	double fourpiG, a, tau, dtau, dtau_old, dx, T00hom;
	int cycle;
	double maxvel[2 + GEVB_MAX_NCDM];                 // by species slot: cdm, baryons, ncdm 0..3 (main.cpp indexes [i+1+baryon_flag])
	Particles_gevolution pcls_cdm, pcls_b;
	Particles_gevolution pcls_ncdm[GEVB_MAX_NCDM];    // main.cpp:219
	double z_switch_deltancdm[GEVB_MAX_NCDM], z_switch_Bncdm[GEVB_MAX_NCDM], z_switch_linearchi, movelimit;   // metadata.hpp:224-238
	int numsteps_ncdm[GEVB_MAX_NCDM];
	Field<Real> phi, source, chi, Sij, Bi;
	Field<Cplx> scalarFT, SijFT, BiFT;
	PlanFFT<Cplx> plan_source, plan_phi, plan_chi, plan_Sij, plan_Bi;
	explicit gevb_sim(gevb_ctx * ctx) : lat(ctx) {}
};


#endif
