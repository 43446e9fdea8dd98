// oracle/latfield2_shim/gsl/gsl_spline.h -- TEST INFRASTRUCTURE
// Stand-in for the GSL spline entry points the reference's IC generator uses (ic_basic.hpp:465-467,708-712:
// gsl_spline_alloc(gsl_interp_cspline, n), gsl_spline_init, gsl_spline_eval, gsl_spline_free, gsl_interp_accel_*).
// GSL is absent in this image.  gsl_interp_cspline is the natural cubic spline (second derivative zero at both
// ends); this header solves the same tridiagonal system, so the interpolant is GSL's up to round-off.
#ifndef GSL_SPLINE_STUB_H
#define GSL_SPLINE_STUB_H
#include <cstddef>
#include <vector>

struct gsl_interp_type { int id; };
static const gsl_interp_type gsl_interp_cspline_obj = {1};
static const gsl_interp_type * const gsl_interp_cspline = &gsl_interp_cspline_obj;

struct gsl_interp_accel { size_t cache; };
inline gsl_interp_accel * gsl_interp_accel_alloc() { gsl_interp_accel * a = new gsl_interp_accel(); a->cache = 0; return a; }
inline void gsl_interp_accel_free(gsl_interp_accel * a) { delete a; }

struct gsl_spline
{
	size_t size;
	double * x, * y;                  // copies of the knots and values (the reference reads spline->x, ->y, ->size directly)
	std::vector<double> c;            // second-derivative coefficients (GSL's c_i: y'' / 2)
};

inline gsl_spline * gsl_spline_alloc(const gsl_interp_type *, size_t size)
{
	gsl_spline * s = new gsl_spline();
	s->size = size;
	s->x = new double[size]; s->y = new double[size];
	return s;
}

inline int gsl_spline_init(gsl_spline * s, const double * xa, const double * ya, size_t size)
{
	if (size != s->size) { delete[] s->x; delete[] s->y; s->x = new double[size]; s->y = new double[size]; s->size = size; }
	for (size_t i = 0; i < size; i++) { s->x[i] = xa[i]; s->y[i] = ya[i]; }
	s->c.assign(size, 0.);
	if (size < 3) return 0;
	// natural boundary conditions: c_0 = c_{n-1} = 0; interior rows h_{i-1} c_{i-1} + 2 (h_{i-1} + h_i) c_i + h_i c_{i+1} = 3 (d_i - d_{i-1})
	const size_t m = size - 2;
	std::vector<double> diag(m), off(m), rhs(m);
	for (size_t i = 0; i < m; i++)
	{
		const double h0 = xa[i + 1] - xa[i], h1 = xa[i + 2] - xa[i + 1];
		const double d0 = (ya[i + 1] - ya[i]) / h0, d1 = (ya[i + 2] - ya[i + 1]) / h1;
		off[i] = h1; diag[i] = 2. * (h0 + h1); rhs[i] = 3. * (d1 - d0);
	}
	// symmetric tridiagonal solve (Thomas algorithm)
	for (size_t i = 1; i < m; i++)
	{
		const double w = off[i - 1] / diag[i - 1];
		diag[i] -= w * off[i - 1];
		rhs[i] -= w * rhs[i - 1];
	}
	std::vector<double> sol(m);
	sol[m - 1] = rhs[m - 1] / diag[m - 1];
	for (size_t i = m - 1; i-- > 0;) sol[i] = (rhs[i] - off[i] * sol[i + 1]) / diag[i];
	for (size_t i = 0; i < m; i++) s->c[i + 1] = sol[i];
	return 0;
}

inline double gsl_spline_eval(const gsl_spline * s, double x, gsl_interp_accel * acc)
{
	const size_t n = s->size;
	size_t i = acc ? acc->cache : 0;
	if (i > n - 2) i = n - 2;
	if (!(s->x[i] <= x && x < s->x[i + 1]))
	{
		size_t lo = 0, hi = n - 1;                       // bisection as gsl_interp_bsearch: largest i with x[i] <= x, clamped to [0, n-2]
		while (hi > lo + 1) { const size_t mid = (lo + hi) / 2; if (s->x[mid] > x) hi = mid; else lo = mid; }
		i = lo;
		if (acc) acc->cache = i;
	}
	const double h = s->x[i + 1] - s->x[i], dy = s->y[i + 1] - s->y[i], delx = x - s->x[i];
	const double b = dy / h - h * (s->c[i + 1] + 2. * s->c[i]) / 3.;
	const double d = (s->c[i + 1] - s->c[i]) / (3. * h);
	return s->y[i] + delx * (b + delx * (s->c[i] + delx * d));
}

inline void gsl_spline_free(gsl_spline * s) { if (s) { delete[] s->x; delete[] s->y; delete s; } }
#endif
