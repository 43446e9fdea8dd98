// background.hpp -- Friedmann background on the host (O(1) work per step)
//
// Follows the reference's background.hpp: FermiDiracIntegral (:19-50), bg_ncdm
// (:69-120), Hconf (:137-140), rungekutta4bg (:167-177), particleHorizon
// (:200-213).  These scalars only feed coefficients of the device kernels.
#ifndef GEVB_HOST_BACKGROUND_HPP
#define GEVB_HOST_BACKGROUND_HPP
#include <cmath>

namespace gevb200 {

#define GEVB_C_SPEED_OF_LIGHT 2997.92458   // speed of light [100 km/s], metadata.hpp:98

#define GEVB_C_PLANCK_LAW 4.48147e-7       // omega_g / (T_cmb [K])^4, metadata.hpp:96
#define GEVB_C_BOLTZMANN_CST 8.61733e-5    // Boltzmann constant [eV/K], metadata.hpp:97
#define GEVB_C_FD_NORM 1.80308535          // Integral[q*q/(exp(q)+1), 0, infinity], metadata.hpp:100
#define GEVB_MAX_NCDM 4                    // MAX_PCL_SPECIES-2, metadata.hpp:48-49

struct cosmology
{
	double Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, w0_fld, wa_fld, Omega_g, Omega_ur, Omega_rad, h;
	int num_ncdm;
	double Omega_ncdm[GEVB_MAX_NCDM], m_ncdm[GEVB_MAX_NCDM], T_ncdm[GEVB_MAX_NCDM];
};

// Integral of q^2 sqrt(q^2 + w) / (e^q + 1) over [0, 24] (background.hpp:19-50).  The reference asks GSL's QNG for
// 1e-7 relative accuracy; the integrand is smooth, so a composite 8-point Gauss-Legendre rule on 48 panels is
// converged to round-off and agrees with any quadrature that meets the reference's request.
inline double FermiDiracIntegral(const double w)
{
	static const double gx[4] = {0.1834346424956498049394761, 0.5255324099163289858177390, 0.7966664774136267395915539, 0.9602898564975362316835609};
	static const double gw[4] = {0.3626837833783619829651504, 0.3137066458778872873379622, 0.2223810344533744705443560, 0.1012285362903762591525314};
	const int panels = 48;
	const double h = 24.0 / panels;
	double sum = 0.;
	for (int p = 0; p < panels; p++)
	{
		const double c = (p + 0.5) * h;
		for (int j = 0; j < 4; j++)
			for (int sgn = -1; sgn <= 1; sgn += 2)
			{
				const double q = c + sgn * 0.5 * h * gx[j];
				sum += gw[j] * q * q * sqrt(q * q + w) / (exp(q) + 1.0);
			}
	}
	return sum * 0.5 * h;
}

inline double bg_ncdm(const double a, const cosmology & cosmo, const int p)                      // background.hpp:69-80
{
	if (p < 0 || p >= cosmo.num_ncdm) return 0;
	double w = a * cosmo.m_ncdm[p] / (pow(cosmo.Omega_g * cosmo.h * cosmo.h / GEVB_C_PLANCK_LAW, 0.25) * cosmo.T_ncdm[p] * GEVB_C_BOLTZMANN_CST);
	w *= w;
	return FermiDiracIntegral(w) * cosmo.Omega_ncdm[p] * pow(cosmo.Omega_g * cosmo.h * cosmo.h / GEVB_C_PLANCK_LAW, 0.25) * cosmo.T_ncdm[p] * GEVB_C_BOLTZMANN_CST / cosmo.m_ncdm[p] / GEVB_C_FD_NORM / a;
}

inline double bg_ncdm(const double a, const cosmology & cosmo)                                   // background.hpp:103-120
{
	double result = 0.0;
	for (int p = 0; p < cosmo.num_ncdm; p++) result += bg_ncdm(a, cosmo, p);
	return result;
}

inline double Hconf(const double a, const double fourpiG, const cosmology & cosmo)
{
	return sqrt((2. * fourpiG / 3.) * (((cosmo.Omega_cdm + cosmo.Omega_b + bg_ncdm(a, cosmo)) / a) + (cosmo.Omega_Lambda * a * a) + (cosmo.Omega_rad / a / a) + (cosmo.Omega_fld * exp(3. * cosmo.wa_fld * (a - 1.)) / pow(a, 1. + 3. * (cosmo.w0_fld + cosmo.wa_fld)))));
}

inline void rungekutta4bg(double & a, const double fourpiG, const cosmology & cosmo, const double dtau)
{
	double k1a, k2a, k3a, k4a;
	k1a = a * Hconf(a, fourpiG, cosmo);
	k2a = (a + k1a * dtau / 2.) * Hconf(a + k1a * dtau / 2., fourpiG, cosmo);
	k3a = (a + k2a * dtau / 2.) * Hconf(a + k2a * dtau / 2., fourpiG, cosmo);
	k4a = (a + k3a * dtau) * Hconf(a + k3a * dtau, fourpiG, cosmo);
	a += dtau * (k1a + 2. * k2a + 2. * k3a + k4a) / 6.;
}

// tau(a) = int_0^sqrt(a) 2 / (s Hconf(s^2)) ds / sqrt(fourpiG); composite Simpson on the
// reference's interval [1e-7 sqrt(a), sqrt(a)] (the reference uses GSL QNG to 1e-7 relative)
inline double particleHorizon(const double a, const double fourpiG, const cosmology & cosmo)
{
	const double lo = sqrt(a) * 1.0e-7, hi = sqrt(a);
	const int n = 4096;
	const double h = (hi - lo) / n;
	double s = 0.;
	for (int i = 0; i <= n; i++)
	{
		const double x = lo + i * h;
		const double f = 2. / (x * Hconf(x * x, 1., cosmo));
		s += (i == 0 || i == n) ? f : ((i & 1) ? 4. * f : 2. * f);
	}
	return s * h / 3. / sqrt(fourpiG);
}

} // namespace gevb200
#endif
