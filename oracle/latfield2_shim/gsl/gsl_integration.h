// oracle/latfield2_shim/gsl/gsl_integration.h -- TEST INFRASTRUCTURE
// Stand-in for the one GSL entry point the reference's background.hpp uses
// (gsl_integration_qng, background.hpp:47,210).  GSL is absent in this image.
// Not the QNG algorithm: an adaptive 15-point Gauss-Kronrod rule driven to a
// tighter tolerance than the reference requests (5e-7 abs / 1e-7 rel), so the
// value agrees with GSL's to better than the reference's own requested error.
// Both sides of every parity comparison use the same host scalars, so exact
// GSL reproduction is not required.
#ifndef GSL_INTEGRATION_STUB_H
#define GSL_INTEGRATION_STUB_H
#include <cmath>
#include <cstddef>

struct gsl_function { double (*function)(double, void *); void * params; };

namespace gslstub {
inline double gk15(const gsl_function * f, double a, double b, double * err)
{
	static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
		0.864864423359769072789712788640926, 0.741531185599394439863864773280788, 0.586087235467691130294144838258730,
		0.405845151377397166906606412076961, 0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
	static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
		0.104790010322250183839876322541518, 0.140653259715525918745189590510238, 0.169004726639267902826583426598550,
		0.190350578064785409913256402421014, 0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
	static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
		0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
	double c = 0.5 * (a + b), h = 0.5 * (b - a);
	double fc = f->function(c, f->params);
	double rk = fc * wgk[7], rg = fc * wg[3];
	for (int j = 0; j < 7; j++)
	{
		double dxx = h * xgk[j];
		double s = f->function(c - dxx, f->params) + f->function(c + dxx, f->params);
		rk += wgk[j] * s;
		if (j % 2 == 1) rg += wg[j / 2] * s;
	}
	*err = fabs((rk - rg) * h);
	return rk * h;
}
inline double adapt(const gsl_function * f, double a, double b, double tol, int depth, size_t * neval)
{
	double err, r = gk15(f, a, b, &err);
	*neval += 15;
	if (err <= tol || depth > 40) return r;
	double m = 0.5 * (a + b);
	return adapt(f, a, m, 0.5 * tol, depth + 1, neval) + adapt(f, m, b, 0.5 * tol, depth + 1, neval);
}
}

inline int gsl_integration_qng(const gsl_function * f, double a, double b, double epsabs, double epsrel, double * result, double * abserr, size_t * neval)
{
	size_t n = 0; double e0;
	double coarse = gslstub::gk15(f, a, b, &e0);
	double tol = 1.0e-13 * fabs(coarse);
	if (tol <= 0.) tol = 1.0e-300;
	*result = gslstub::adapt(f, a, b, tol, 0, &n);
	*abserr = tol; *neval = n;
	(void) epsabs; (void) epsrel;
	return 0;
}
#endif
