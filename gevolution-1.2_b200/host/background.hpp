// background.hpp -- Friedmann background on the host (O(1) work per step)
//
// Follows the reference's background.hpp: Hconf (:137-140), rungekutta4bg
// (:167-177), particleHorizon (:200-213) for cosmologies without ncdm species
// (bg_ncdm == 0).  These scalars only feed coefficients of the device kernels.
#ifndef GEVB_HOST_BACKGROUND_HPP
#define GEVB_HOST_BACKGROUND_HPP
#include <cmath>

namespace gevb200 {

#define GEVB_C_SPEED_OF_LIGHT 2997.92458   // speed of light [100 km/s], metadata.hpp:98

struct cosmology
{
	double Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, w0_fld, wa_fld, Omega_g, Omega_ur, Omega_rad, h;
};

inline double bg_ncdm(const double, const cosmology &) { return 0.; }

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
