// ic_basic.cpp -- "IC generator = basic" (generateIC_basic, ic_basic.hpp:1626-2259) on the host, over the device calls
//
// What runs where.  Host: the Gadget-2 particle template (loadHomogeneousTemplate :191-330), the transfer-function
// table and its splines (loadTransferFunctions :494-717, gsl_interp_cspline = natural cubic spline), the 27-site CIC
// convolution kernel (generateCICKernel :737-1052), the Gaussian realisation in Fourier space (generateDisplacementField
// :1090-1379 with the Threefry-4x64-20 counter stream of prng_engine.hpp in sitmo's output order) and the template
// tiling (initializeParticlePositions :1400-1434).  Device, through the calls of the hot path: every FFT, the particle
// displacement and initial momenta (the displace_pcls_ic_basic / initialize_q_ic_basic callbacks), and the initial
// phi, chi, B from the T0i / Tij projections and their Fourier-space projections.
//
// The realisation is organised by rows of Fourier space instead of the reference's four sequential sweeps: the
// reference gives every row (ky, kz) of every quadrant its own stretch of 65536 draws of the stream (HUGE_SKIP), so a
// row's first draw sits at a position that follows from (quadrant, ky, kz) alone and rows can be filled independently
// -- here by a pool of host threads, on any ky-slab of the lattice.  The arithmetic of a mode is the reference's,
// including its single-precision intermediates (ic_basic.hpp:1098,1161-1168), so the same seed gives the same field.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <thread>
#include "sim_internal.hpp"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Threefry-4x64 with 20 rounds (Salmon et al. 2011), key = (seed, 0, 0, 0), counter = (block, 0, 0, 0), read as the
// stream of 32-bit words sitmo::prng_engine delivers: eight words per block, low half of a 64-bit lane first
// (prng_engine.hpp:205-222); discard() of the reference is plain arithmetic on the word position (:230-246).
class CounterStream
{
	uint64_t ks_[5];
	uint64_t block_, out_[4];
	bool have_;
	static uint64_t rotl(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }
	void cipher(uint64_t block)
	{
		static const int rot[8][2] = {{14, 16}, {52, 57}, {23, 40}, {5, 37}, {25, 33}, {46, 12}, {58, 22}, {32, 32}};
		uint64_t x[4] = {block, 0, 0, 0};
		for (int round = 0; round < 20; round++)
		{
			if ((round & 3) == 0)
			{
				const int s = round >> 2;
				for (int i = 0; i < 4; i++) x[i] += ks_[(s + i) % 5];
				x[3] += (uint64_t) s;
			}
			const int a = (round & 1) ? 3 : 1, b = (round & 1) ? 1 : 3;       // even rounds mix (0,1),(2,3); odd rounds (0,3),(2,1)
			x[0] += x[a]; x[a] = rotl(x[a], rot[round & 7][0]) ^ x[0];
			x[2] += x[b]; x[b] = rotl(x[b], rot[round & 7][1]) ^ x[2];
		}
		for (int i = 0; i < 4; i++) out_[i] = x[i] + ks_[(5 + i) % 5];
		out_[3] += 5;
		block_ = block; have_ = true;
	}
public:
	explicit CounterStream(uint32_t seed) : block_(0), have_(false)
	{
		ks_[0] = seed; ks_[1] = ks_[2] = ks_[3] = 0;
		ks_[4] = 0x1BD11BDAA9FC1A22ull ^ ks_[0] ^ ks_[1] ^ ks_[2] ^ ks_[3];
	}
	uint32_t word(uint64_t pos)
	{
		if (!have_ || (pos >> 3) != block_) cipher(pos >> 3);
		const uint64_t lane = out_[(pos & 7) >> 1];
		return (pos & 1) ? (uint32_t) (lane >> 32) : (uint32_t) (lane & 0xFFFFFFFFull);
	}
};

// ---------------------------------------------------------------------------------------------------------------
// natural cubic spline (gsl_interp_cspline as used at ic_basic.hpp:465-467,708-712): second derivative zero at both ends
struct Spline
{
	std::vector<double> x, y, c;                                   // knots, values, c_i = y''(x_i) / 2
	void init(const double * xa, const double * ya, size_t n)
	{
		x.assign(xa, xa + n); y.assign(ya, ya + n); c.assign(n, 0.);
		if (n < 3) return;
		// rows i = 1 .. n-2:  h_{i-1} c_{i-1} + 2 (h_{i-1} + h_i) c_i + h_i c_{i+1} = 3 ((y_{i+1} - y_i) / h_i - (y_i - y_{i-1}) / h_{i-1})
		const size_t m = n - 2;
		std::vector<double> diag(m), off(m), rhs(m);
		for (size_t i = 0; i < m; i++)
		{
			const double hl = x[i + 1] - x[i], hr = x[i + 2] - x[i + 1];
			diag[i] = 2. * (hl + hr); off[i] = hr;
			rhs[i] = 3. * ((y[i + 2] - y[i + 1]) / hr - (y[i + 1] - y[i]) / hl);
		}
		for (size_t i = 1; i < m; i++)                             // forward elimination of the symmetric tridiagonal system
		{
			const double f = off[i - 1] / diag[i - 1];
			diag[i] -= f * off[i - 1]; rhs[i] -= f * rhs[i - 1];
		}
		c[m] = rhs[m - 1] / diag[m - 1];
		for (size_t i = m - 1; i-- > 0;) c[i + 1] = (rhs[i] - off[i] * c[i + 2]) / diag[i];
	}
	double eval(double xv, size_t & hint) const
	{
		const size_t n = x.size();
		size_t i = hint > n - 2 ? n - 2 : hint;
		if (!(x[i] <= xv && xv < x[i + 1]))
		{
			size_t lo = 0, hi = n - 1;                             // largest i with x[i] <= xv, clamped into [0, n-2]
			while (hi > lo + 1) { const size_t mid = (lo + hi) / 2; if (x[mid] > xv) hi = mid; else lo = mid; }
			i = hint = lo;
		}
		const double h = x[i + 1] - x[i], t = xv - x[i];
		const double b = (y[i + 1] - y[i]) / h - h * (c[i + 1] + 2. * c[i]) / 3.;
		const double d = (c[i + 1] - c[i]) / (3. * h);
		return y[i] + t * (b + t * (c[i] + t * d));
	}
};

// ---------------------------------------------------------------------------------------------------------------
// generateCICKernel (ic_basic.hpp:737-1052).  The kernel lives on the 3 x 3 x 3 sites around the origin; K[(dz+1)*9 +
// (dy+1)*3 + (dx+1)] collects site (dx, dy, dz) mod N.  The reference unrolls eight octants times three directions; the
// same sums are formed here from one description: in direction d the particle's CIC cloud (weights w / 1-w along the two
// other axes) is differenced along d with the stencil (+q at 0, -1 at +s_d, -1 at -s_d when w_d > 0.9, q = 2 then).
// Operand types follow the reference literally -- w is float, (1. - w) is double, a product of two floats stays float --
// because the sums are compared to round-off.
struct Num { double v; bool f32; };
inline Num times(Num a, Num b)
{
	if (a.f32 && b.f32) { const float p = (float) a.v * (float) b.v; return Num{(double) p, true}; }
	return Num{a.v * b.v, false};
}

void cic_kernel(int N, long numpcl, const float * pcldata, int numtile, double * K)
{
	for (int i = 0; i < 27; i++) K[i] = 0.;
	const long linesize = N;
	double renorm = (double) (linesize * linesize);
	auto at = [&](int dx, int dy, int dz) -> double & { return K[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)]; };
	if (numpcl == 0 || pcldata == NULL)                                       // standard kernel: the 7-point Laplacian (:750-773)
	{
		at(0, 0, 0) = 6. * renorm;
		at(1, 0, 0) = at(-1, 0, 0) = at(0, 1, 0) = at(0, -1, 0) = at(0, 0, 1) = at(0, 0, -1) = -renorm;
		return;
	}
	renorm /= (double) (numpcl * (long) numtile * (long) numtile * (long) numtile) / (double) (linesize * linesize * linesize);   // :777
	for (long i = 0; i < numpcl; i++)
	{
		for (int oct = 0; oct < 8; oct++)
		{
			// side of the particle along x, y, z in this octant: +1 measures from the lower cell face, -1 from the upper one
			static const int side[8][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}, {1, -1, -1}};
			float w[3];
			bool inside = true;
			for (int l = 0; l < 3; l++)
			{
				const float p = pcldata[3 * i + l];
				if (side[oct][l] > 0)
				{
					const float t = (float) linesize * p / (float) numtile;      // linesize * pcldata[] / numtile in single precision (:785)
					if (t >= 1.) inside = false;
					w[l] = (float) (1. - t);
				}
				else
				{
					const double t = (double) linesize * (1. - p) / (double) numtile;   // (1. - pcldata[]) promotes the product to double (:799)
					if (t >= 1.) inside = false;
					w[l] = (float) (1. - t);
				}
			}
			if (!inside) continue;
			const int * sg = side[oct];
			for (int d = 0; d < 3; d++)
			{
				const int u = d == 0 ? 1 : 0, v = d == 2 ? 1 : 2;                // the two other axes, in increasing order
				const float ww = (float) ((double) (w[u] * w[v]) * renorm);      // float ww = w_u * w_v * renorm (:907,:951,:1001)
				const bool wide = w[d] > 0.9;
				const Num q = {wide ? 2. : 1., true};
				for (int ou = 0; ou < 2; ou++)
					for (int ov = 0; ov < 2; ov++)
					{
						const Num fu = ou == 0 ? Num{(double) w[u], true} : Num{1. - (double) w[u], false};
						const Num fv = ov == 0 ? Num{(double) w[v], true} : Num{1. - (double) w[v], false};
						const Num term = times(times(Num{(double) ww, true}, fu), fv);
						int off[3];
						off[u] = ou * sg[u]; off[v] = ov * sg[v];
						off[d] = 0; at(off[0], off[1], off[2]) += times(term, q).v;
						off[d] = sg[d]; at(off[0], off[1], off[2]) -= term.v;
						if (wide) { off[d] = -sg[d]; at(off[0], off[1], off[2]) -= term.v; }
					}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// generateDisplacementField (ic_basic.hpp:1090-1379), one row (ky, kz) of Fourier space at a time.
struct Realisation
{
	int N, kmax, ksphere;
	double coeff;
	const Spline * pk;
	uint32_t seed;
	std::vector<float> sinc;
	Realisation(int n, double co, const Spline * sp, uint32_t sd, int ks, int deconvolve_f) : N(n), kmax(n / 2 - 1), ksphere(ks), coeff(co), pk(sp), seed(sd), sinc((size_t) n)
	{
		sinc[0] = 1.;
		for (int i = 1; i < N; i++)
			sinc[i] = deconvolve_f == 1 ? (float) (std::sin(M_PI * (float) i / (float) N) * (float) N / (M_PI * (float) i)) : 1.f;   // :1106-1117
	}
	// value of one mode given |k| components (a, b, c) = (kx, ky or N-ky, kz or N-kz); `pos` is the stream position of the next draw
	bool mode(CounterStream & rng, uint64_t & pos, int a, int b, int c, bool conj, size_t & hint, double * inout) const
	{
		float k2 = (float) (a * a) + (float) (b * b) + (float) (c * c);
		if (a >= kmax || b >= kmax || c >= kmax || (k2 >= kmax * kmax && ksphere > 0)) { inout[0] = inout[1] = 0.; return false; }   // :1150,:1154
		const float s = sinc[a] * sinc[b] * sinc[c];
		k2 *= 4. * M_PI * M_PI;
		float r1, r2;
		do { r1 = (float) rng.word(pos++) / (float) 0xFFFFFFFFu; } while (r1 == 0.);                   // :1160-1165
		r2 = (float) rng.word(pos++) / (float) 0xFFFFFFFFu;
		// Cplx(cos, +-sin) * (1 + 7.5 coeff / k2) / potFT(k) * sqrt(-2 log r1) * P(sqrt(k2)) * s   (:1168; conjugated on the kx = 0 planes of the upper half :1278)
		const double ph = 2. * M_PI * r2;
		double re = std::cos(ph), im = conj ? -std::sin(ph) : std::sin(ph);
		const double boost = 1. + 7.5 * coeff / k2;
		re *= boost; im *= boost;
		const double kr = inout[0], ki = inout[1], den = kr * kr + ki * ki;                            // complex division by the kernel's transform
		double qr = (re * kr + im * ki) / den, qi = (im * kr - re * ki) / den;
		// the three real factors multiply the complex number one after the other, as the reference's expression does
		const double gauss = std::sqrt(-2. * std::log(r1)), amp = pk->eval(std::sqrt(k2), hint);
		qr *= gauss; qi *= gauss;
		qr *= amp; qi *= amp;
		inout[0] = qr * s; inout[1] = qi * s;
		return true;
	}
	// row of nh = N/2 + 1 complex numbers (kx = 0 .. N/2) at (ky, kz)
	void row(int ky, int kz, double * data) const
	{
		static const uint64_t H = 65536;                                                               // HUGE_SKIP (ic_basic.hpp:31-33)
		const int nh = N / 2 + 1;
		const bool up_y = ky > N / 2, up_z = kz > N / 2;
		const int b = up_y ? N - ky : ky, c = up_z ? N - kz : kz;
		const uint64_t quadrant = (up_y ? 1 : 0) + (up_z ? 2 : 0);
		CounterStream rng(seed);
		size_t hint = 0;
		uint64_t pos = ((quadrant * H + (uint64_t) c) * H + (uint64_t) b) * H;                         // first draw of the row (:1131,:1188,:1221,:1291)
		if (!up_z)
		{
			// lower half in kz: the row is drawn from kx = 0 on (the origin itself is set to zero without a draw, :1133-1137)
			int kx = 0;
			if (ky == 0 && kz == 0) { data[0] = data[1] = 0.; kx = 1; }
			for (; kx < nh; kx++) mode(rng, pos, kx, b, c, false, hint, data + 2 * kx);
		}
		else
		{
			// upper half in kz: kx = 1 .. N/2 from this quadrant's stream (:1225,:1295); the kx = 0 plane repeats, conjugated, the
			// first draws of the row (b, c) of the quadrant two below (:1249-1283, :1319-1353)
			for (int kx = 1; kx < nh; kx++) mode(rng, pos, kx, b, c, false, hint, data + 2 * kx);
			uint64_t pos0 = (((quadrant - 2) * H + (uint64_t) c) * H + (uint64_t) b) * H;
			if (up_y) pos0 = ((2 * H + (uint64_t) c) * H + (uint64_t) b) * H;                          // :1320: huge+huge, not the mirror quadrant
			mode(rng, pos0, 0, b, c, true, hint, data);
		}
	}
};

// the realisation on the rows [ky0, ky0 + nky) x [0, N) of a host array laid out [kz][ky][kx] (slab == false) or [ky - ky0][kz][kx]
void realise(int N, int ky0, int nky, bool slab, double * potFT, double coeff, const Spline & pk, uint32_t seed, int ksphere, int deconvolve_f)
{
	const Realisation R(N, coeff, &pk, seed, ksphere, deconvolve_f);
	const int nh = N / 2 + 1;
	unsigned nthreads = std::thread::hardware_concurrency();
	if (nthreads < 1) nthreads = 1;
	if (nthreads > 32) nthreads = 32;
	if ((long) N * nky < 4096) nthreads = 1;
	auto work = [&](unsigned t)
	{
		for (int kz = (int) t; kz < N; kz += (int) nthreads)
			for (int j = 0; j < nky; j++)
			{
				const size_t r = slab ? (size_t) j * N + kz : (size_t) kz * N + (ky0 + j);
				R.row(ky0 + j, kz, potFT + 2 * r * nh);
			}
	};
	if (nthreads == 1) { work(0); return; }
	std::vector<std::thread> pool;
	for (unsigned t = 0; t < nthreads; t++) pool.emplace_back(work, t);
	for (std::thread & th : pool) th.join();
}

// ---------------------------------------------------------------------------------------------------------------
// loadHomogeneousTemplate (ic_basic.hpp:191-330): Gadget-2 file, positions of particle type 1 in units of the file's box
int load_template(const char * filename, long & numpart, std::vector<float> & data)
{
	struct Header { uint32_t npart[6]; double mass[6]; double time, redshift; int32_t flag_sfr, flag_feedback; uint32_t npartTotal[6]; int32_t flag_cooling, num_files;
	                double BoxSize, Omega0, OmegaLambda, HubbleParam; int32_t flag_age, flag_metals; uint32_t npartTotalHW[6]; char fill[64]; } hdr;   // metadata.hpp:152-171
	static_assert(sizeof(Header) == 256, "gadget2 header must be 256 bytes");
	FILE * f = std::fopen(filename, "rb");
	if (f == NULL) { std::fprintf(stderr, " error in loadHomogeneousTemplate! Unable to open template file %s.\n", filename); return 1; }
	int32_t b1 = 0, b2 = 0;
	bool ok = std::fread(&b1, 4, 1, f) == 1 && b1 == (int32_t) sizeof(hdr) && std::fread(&hdr, sizeof(hdr), 1, f) == 1 && std::fread(&b2, 4, 1, f) == 1 && b1 == b2;
	if (!ok) { std::fprintf(stderr, " error in loadHomogeneousTemplate! Unknown template file format - header not recognized.\n"); std::fclose(f); return 1; }
	if (hdr.num_files != 1 || !(hdr.BoxSize > 0.) || hdr.npart[1] == 0)
	{
		std::fprintf(stderr, " error in loadHomogeneousTemplate! Unsupported template (files %d, BoxSize %g, particles %u).\n", hdr.num_files, hdr.BoxSize, hdr.npart[1]);
		std::fclose(f); return 1;
	}
	data.resize(3 * (size_t) hdr.npart[1]);
	ok = std::fread(&b1, 4, 1, f) == 1 && (hdr.npart[0] == 0 || std::fseek(f, (long) (3 * sizeof(float) * hdr.npart[0]), SEEK_CUR) == 0)
		&& std::fread(data.data(), sizeof(float), data.size(), f) == data.size();
	for (int i = 2; i < 6 && ok; i++) if (hdr.npart[i] > 0) ok = std::fseek(f, (long) (3 * sizeof(float) * hdr.npart[i]), SEEK_CUR) == 0;
	ok = ok && std::fread(&b2, 4, 1, f) == 1 && b1 == b2;
	std::fclose(f);
	if (!ok) { std::fprintf(stderr, " error in loadHomogeneousTemplate! Unable to read particle data.\n"); return 1; }
	for (float & v : data)
	{
		v /= hdr.BoxSize;                                                                              // :303 (float /= double)
		if (v < 0. || v > 1.) { std::fprintf(stderr, " error in loadHomogeneousTemplate! Particle data corrupted.\n"); return 1; }
	}
	numpart = (long) hdr.npart[1];
	return 0;
}

// loadTransferFunctions (ic_basic.hpp:494-717): columns "k", "d_<qname>", "t_<qname>" of a CLASS transfer-function table,
// identified by the header line "# 1:k (h/Mpc)  2:d_g ..."; k in units of the box, theta in units of box / h
int load_transfer(const char * filename, const char * qname, double boxsize, double h, std::vector<double> & k, std::vector<double> & td, std::vector<double> & tt)
{
	FILE * f = std::fopen(filename, "r");
	if (f == NULL) { std::fprintf(stderr, " error in loadTransferFunctions! Unable to open file %s.\n", filename); return 1; }
	std::vector<std::string> lines;
	{
		std::string cur;
		int ch;
		while ((ch = std::fgetc(f)) != EOF) { if (ch == '\n') { lines.push_back(cur); cur.clear(); } else cur += (char) ch; }
		if (!cur.empty()) lines.push_back(cur);
		std::fclose(f);
	}
	int kcol = -1, dcol = -1, tcol = -1;
	const size_t qlen = std::strlen(qname);
	for (const std::string & line : lines)
	{
		int col = 0;
		for (size_t p = line.find(':'); p != std::string::npos; p = line.find(':', p + 1), col++)
		{
			const char * q = line.c_str() + p + 1;
			if (*q == 'k') kcol = col;
			else if (*q == 'd' && std::strncmp(q + 2, qname, qlen) == 0) dcol = col;
			else if (*q == 't' && std::strncmp(q + 2, qname, qlen) == 0) tcol = col;
		}
		if (kcol >= 0 && dcol >= 0 && tcol >= 0) break;
	}
	if (kcol < 0 || dcol < 0 || tcol < 0) { std::fprintf(stderr, " error in loadTransferFunctions! Unable to identify requested columns (%s)!\n", qname); return 1; }
	k.clear(); td.clear(); tt.clear();
	for (const std::string & line : lines)
	{
		if (line.empty() || line[0] == '#') continue;
		std::vector<double> vals;
		const char * q = line.c_str();
		char * end;
		for (double v = std::strtod(q, &end); end != q; v = std::strtod(q, &end)) { vals.push_back(v); q = end; }
		const int need = std::max(kcol, std::max(dcol, tcol));
		if ((int) vals.size() <= need) continue;
		if (vals[kcol] < 0.) { std::fprintf(stderr, " error in loadTransferFunctions! Negative k-value encountered.\n"); return 1; }
		if (!k.empty() && k.back() >= vals[kcol] * boxsize) { std::fprintf(stderr, " error in loadTransferFunctions! k-values are not strictly ordered.\n"); return 1; }
		k.push_back(vals[kcol] * boxsize);                                                             // :666-668
		td.push_back(vals[dcol]);
		tt.push_back(vals[tcol] * boxsize / h);
	}
	if (k.size() < 2) { std::fprintf(stderr, " error in loadTransferFunctions! No valid data found in file %s.\n", filename); return 1; }
	return 0;
}

inline double Pk_primordial(double k, const gevb_settings & st) { return st.A_s * std::pow(k / st.k_pivot, st.n_s - 1.); }   // ic_basic.hpp:170-173

// ---------------------------------------------------------------------------------------------------------------
// device-side helpers of the driver
struct Generator
{
	gevb_sim * s;
	const gevb_settings & st;
	int N, rank, nranks, z0, nzl, ky0, nkyl;
	std::vector<double> hostFT;                                    // scalarFT of this rank on the host
	Generator(gevb_sim * sim, const gevb_settings & settings) : s(sim), st(settings)
	{
		gevb_ctx_ranks(s->lat.ctx(), &rank, &nranks);
		gevb_ctx_geometry(s->lat.ctx(), &N, &z0, &nzl, &ky0, &nkyl);
		hostFT.resize(2 * (size_t) (N / 2 + 1) * N * (nranks == 1 ? N : nkyl));
	}
	// the kernel of generateCICKernel written into a zeroed field
	void set_kernel(Field<Real> & fld, long numpcl, const float * pcldata, int numtile)
	{
		double K[27];
		cic_kernel(N, numpcl, pcldata, numtile, K);
		projection_init(&fld);
		// the 27 offsets are folded onto lattice sites first (they coincide for N < 3), in the order the reference touches them
		std::vector<int> xyz;
		std::vector<double> val;
		for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++)
		{
			const double v = K[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)];
			if (v == 0.) continue;
			const int c[3] = {(N + dx) % N, (N + dy) % N, (N + dz) % N};
			size_t j = 0;
			for (; j < val.size(); j++) if (xyz[3 * j] == c[0] && xyz[3 * j + 1] == c[1] && xyz[3 * j + 2] == c[2]) break;
			if (j == val.size()) { xyz.insert(xyz.end(), c, c + 3); val.push_back(v); } else val[j] += v;
		}
		check(gevb_field_set_sites(fld.handle(), 0, (int) val.size(), xyz.data(), val.data()), "generateCICKernel");
	}
	// generateDisplacementField on scalarFT (which holds the transform of the kernel)
	void displacement_field(double coeff, const Spline & pk, int deconvolve_f = 1)
	{
		s->scalarFT.download(hostFT.data());
		realise(N, nranks == 1 ? 0 : ky0, nranks == 1 ? N : nkyl, nranks > 1, hostFT.data(), coeff, pk, (uint32_t) st.seed, st.ksphere, deconvolve_f);
		s->scalarFT.upload(hostFT.data());
	}
	// initializeParticlePositions (ic_basic.hpp:1400-1434): the template repeated numtile^3 times; a rank keeps the tiles that
	// can reach its slab and the container files what falls inside
	void tile_template(Particles_gevolution & pcls, long numpart, const float * partdata, int numtile)
	{
		const long zt0 = ((long) z0 * numtile) / N, zt1 = std::min((long) numtile - 1, ((long) (z0 + nzl) * numtile) / N);
		const size_t per_layer = (size_t) numpart * numtile * numtile;
		std::vector<int64_t> id(per_layer);
		std::vector<double> pos(3 * per_layer), vel(3 * per_layer, 0.);
		for (long ztile = zt0; ztile <= zt1; ztile++)
		{
			size_t n = 0;
			for (long ytile = 0; ytile < numtile; ytile++)
				for (long xtile = 0; xtile < numtile; xtile++)
					for (long i = 0; i < numpart; i++, n++)
					{
						pos[3 * n] = ((double) xtile + partdata[3 * i]) / (double) numtile;            // :1422-1424
						pos[3 * n + 1] = ((double) ytile + partdata[3 * i + 1]) / (double) numtile;
						pos[3 * n + 2] = ((double) ztile + partdata[3 * i + 2]) / (double) numtile;
						id[n] = i + numpart * (xtile + (long) numtile * (ytile + (long) numtile * ztile));   // :1426
					}
			pcls.addParticles_global((int64_t) n, id.data(), pos.data(), vel.data());
		}
	}
};

// product M_PI sqrt(P_prim(k h / box) / k) / k that turns a transfer function into a spline of the realisation (:1721 ...)
inline double prim(double x, const gevb_settings & st, double h) { return M_PI * std::sqrt(Pk_primordial(x * h / st.boxsize, st) / x) / x; }

int generate(gevb_sim * s, const gevb_settings & st)
{
	Generator G(s, st);
	cosmology & cosmo = s->cosmo;
	const double fourpiG = s->fourpiG, a = 1. / (1. + st.z_in), h = cosmo.h;
	const int gr = st.gr_flag;
	int baryon_flag = st.baryon_flag;
	long numpcl0 = 0, numpcl1 = 0;
	std::vector<float> pcldata;
	Field<Real> * ic_fields[2] = {&s->chi, &s->phi};                                                   // :1656-1657
	double max_displacement = 0.;
	int reduce_max = MAX;                                                                              // i = MAX (:1994)

	if (load_template(st.template_file[0], numpcl0, pcldata) != 0) return 1;                          // :1665
	if (st.correct_displacement) G.set_kernel(s->source, numpcl0, pcldata.data(), st.tiling[0]);       // :1673-1676
	else G.set_kernel(s->source, 0, NULL, 1);
	s->plan_source.execute(FFT_FORWARD);                                                               // :1678

	// ---- splines from the transfer functions (:1704-1980)
	std::vector<double> k, d1, t1, d2, t2, kk, dd, tt;
	if (load_transfer(st.tk_file, "tot", st.boxsize, h, k, d1, t1) != 0) return 1;                     // :1713
	const size_t n = k.size();
	std::vector<double> temp1(n), temp2(n);
	const double H = Hconf(a, fourpiG, cosmo);
	const double rescale = 3. * H * H * H * (1. + 0.5 * H * H * ((1. / Hconf(0.98 * a, fourpiG, cosmo) / Hconf(0.98 * a, fourpiG, cosmo)) - (8. / Hconf(0.99 * a, fourpiG, cosmo) / Hconf(0.99 * a, fourpiG, cosmo))
		+ (8. / Hconf(1.01 * a, fourpiG, cosmo) / Hconf(1.01 * a, fourpiG, cosmo)) - (1. / Hconf(1.02 * a, fourpiG, cosmo) / Hconf(1.02 * a, fourpiG, cosmo))) / 0.12);   // :1724
	for (size_t i = 0; i < n; i++)                                                                     // construct phi (:1725-1726)
		temp1[i] = (1.5 * (H * H - Hconf(1., fourpiG, cosmo) * Hconf(1., fourpiG, cosmo) * a * a * cosmo.Omega_Lambda) * d1[i] + rescale * t1[i] / k[i] / k[i]) * M_PI * std::sqrt(Pk_primordial(k[i] * h / st.boxsize, st) / k[i]) / k[i];
	Spline pkspline, nbspline, tk_d1, tk_t1, tk_d2, tk_t2;
	if (gr == 0)
	{
		for (size_t i = 0; i < n; i++)                                                                 // gauge correction for N-body gauge (:1730-1731)
			temp2[i] = -3. * H * M_PI * t1[i] * std::sqrt(Pk_primordial(k[i] * h / st.boxsize, st) / k[i]) / k[i] / k[i] / k[i];
		nbspline.init(k.data(), temp2.data(), n);
	}
	pkspline.init(k.data(), temp1.data(), n);
	if (load_transfer(st.tk_file, "cdm", st.boxsize, h, kk, d1, t1) != 0) return 1;                    // :1781
	if (baryon_flag > 0)
	{
		if (load_transfer(st.tk_file, "b", st.boxsize, h, kk, d2, t2) != 0) return 1;                  // :1797
		if (d2.size() != d1.size()) { std::fprintf(stderr, " error: baryon transfer function line number mismatch!\n"); return 1; }
	}
	// displacement: -3 phi / k^2 - delta (GR) or N-body gauge shift - delta (Newton); velocity potential: -a theta   (:1812 ...)
	auto gauge = [&](size_t i) { return gr > 0 ? -3. * pkspline.y[i] / pkspline.x[i] / pkspline.x[i] : nbspline.y[i]; };
	const double Ocb = cosmo.Omega_cdm + cosmo.Omega_b;
	if (baryon_flag == 2)                                                                              // blend: weighted average (:1808-1842)
	{
		for (size_t i = 0; i < n; i++)
		{
			temp1[i] = gauge(i) - ((cosmo.Omega_cdm * d1[i] + cosmo.Omega_b * d2[i]) / Ocb) * prim(k[i], st, h);
			temp2[i] = -a * ((cosmo.Omega_cdm * t1[i] + cosmo.Omega_b * t2[i]) / Ocb) * prim(k[i], st, h);
		}
		tk_d1.init(k.data(), temp1.data(), n); tk_t1.init(k.data(), temp2.data(), n);
	}
	if (baryon_flag == 3)                                                                              // hybrid (:1844-1917)
	{
		const bool many_b = 8. * cosmo.Omega_b / Ocb > 1.;
		for (size_t i = 0; i < n; i++)
		{
			const double wd = many_b ? (8. * cosmo.Omega_cdm * d1[i] + (7. * cosmo.Omega_b - cosmo.Omega_cdm) * d2[i]) / Ocb / 7. : ((cosmo.Omega_cdm - 7. * cosmo.Omega_b) * d1[i] + 8. * cosmo.Omega_b * d2[i]) / Ocb;
			const double wt = many_b ? (8. * cosmo.Omega_cdm * t1[i] + (7. * cosmo.Omega_b - cosmo.Omega_cdm) * t2[i]) / Ocb / 7. : ((cosmo.Omega_cdm - 7. * cosmo.Omega_b) * t1[i] + 8. * cosmo.Omega_b * t2[i]) / Ocb;
			temp1[i] = gauge(i) - wd * prim(k[i], st, h);
			temp2[i] = -a * wt * prim(k[i], st, h);
		}
		if (many_b) { tk_d1.init(k.data(), temp1.data(), n); tk_t1.init(k.data(), temp2.data(), n); }
		else { tk_d2.init(k.data(), temp1.data(), n); tk_t2.init(k.data(), temp2.data(), n); }
	}
	if (baryon_flag == 1 || (baryon_flag == 3 && 8. * cosmo.Omega_b / Ocb > 1.))                       // baryonic displacement & velocity (:1919-1946)
	{
		for (size_t i = 0; i < n; i++)
		{
			temp1[i] = gauge(i) - d2[i] * prim(k[i], st, h);
			temp2[i] = -a * t2[i] * prim(k[i], st, h);
		}
		tk_d2.init(k.data(), temp1.data(), n); tk_t2.init(k.data(), temp2.data(), n);
	}
	if (baryon_flag < 2 || (baryon_flag == 3 && 8. * cosmo.Omega_b / Ocb <= 1.))                       // CDM displacement & velocity (:1948-1975)
	{
		for (size_t i = 0; i < n; i++)
		{
			temp1[i] = gauge(i) - d1[i] * prim(k[i], st, h);
			temp2[i] = -a * t1[i] * prim(k[i], st, h);
		}
		tk_d1.init(k.data(), temp1.data(), n); tk_t1.init(k.data(), temp2.data(), n);
	}
	if ((baryon_flag == 1 && !st.correct_displacement) || baryon_flag == 3)                            // :1977-1984
	{
		G.displacement_field(0., tk_d2);
		s->plan_phi.execute(FFT_BACKWARD);
		s->phi.updateHalo();                                                                           // phi now contains the baryonic displacement
		s->plan_source.execute(FFT_FORWARD);
	}
	G.displacement_field(0., tk_d1);                                                                   // :1986
	s->plan_chi.execute(FFT_BACKWARD);                                                                 // :1990
	s->chi.updateHalo();                                                                               // chi now contains the CDM displacement

	// ---- CDM particles (:1993-2010)
	part_simple_info info;
	std::strcpy(info.type_name, "part_simple");
	const long tile3_0 = (long) st.tiling[0] * (long) st.tiling[0] * (long) st.tiling[0];
	info.mass = (baryon_flag == 1 ? cosmo.Omega_cdm : cosmo.Omega_cdm + cosmo.Omega_b) / (Real) (numpcl0 * tile3_0);
	info.relativistic = false;
	s->pcls_cdm.initialize(info, &s->lat);
	G.tile_template(s->pcls_cdm, numpcl0, pcldata.data(), st.tiling[0]);
	if (baryon_flag == 3) s->pcls_cdm.moveParticles(displace_pcls_ic_basic, 1., ic_fields, 2, NULL, &max_displacement, &reduce_max, 1);
	else { Field<Real> * f1[1] = {&s->chi}; s->pcls_cdm.moveParticles(displace_pcls_ic_basic, 1., f1, 1, NULL, &max_displacement, &reduce_max, 1); }

	// ---- baryon particles (:2004-2040)
	if (baryon_flag == 1)
	{
		if (load_template(st.template_file[1], numpcl1, pcldata) != 0) return 1;
		if (st.correct_displacement)
		{
			G.set_kernel(s->phi, numpcl1, pcldata.data(), st.tiling[1]);
			s->plan_phi.execute(FFT_FORWARD);
			G.displacement_field(0., tk_d2);
			s->plan_phi.execute(FFT_BACKWARD);
			s->phi.updateHalo();
		}
		const long tile3_1 = (long) st.tiling[1] * (long) st.tiling[1] * (long) st.tiling[1];
		info.mass = cosmo.Omega_b / (Real) (numpcl1 * tile3_1);
		s->pcls_b.initialize(info, &s->lat);
		s->baryon_flag = 1;
		G.tile_template(s->pcls_b, numpcl1, pcldata.data(), st.tiling[1]);
		Field<Real> * f1[1] = {&s->phi};
		s->pcls_b.moveParticles(displace_pcls_ic_basic, 1., f1, 1, NULL, &max_displacement, &reduce_max, 1);
	}

	// ---- velocities from the transfer functions (:2042-2070)
	if (st.correct_displacement) G.set_kernel(s->source, 0, NULL, 1);
	s->plan_source.execute(FFT_FORWARD);
	if (baryon_flag == 1 || baryon_flag == 3)
	{
		G.displacement_field(0., tk_t2, 0);
		s->plan_phi.execute(FFT_BACKWARD);
		s->phi.updateHalo();                                                                           // phi now contains the baryonic velocity potential
		s->plan_source.execute(FFT_FORWARD);
	}
	G.displacement_field(0., tk_t1, 0);
	s->plan_chi.execute(FFT_BACKWARD);
	s->chi.updateHalo();                                                                               // chi now contains the CDM velocity potential
	double maxvel[2] = {0., 0.};
	if (baryon_flag == 3) maxvel[0] = s->pcls_cdm.updateVel(initialize_q_ic_basic, 1., ic_fields, 2) / a;
	else { Field<Real> * f1[1] = {&s->chi}; maxvel[0] = s->pcls_cdm.updateVel(initialize_q_ic_basic, 1., f1, 1) / a; }
	if (baryon_flag == 1) { Field<Real> * f1[1] = {&s->phi}; maxvel[1] = s->pcls_b.updateVel(initialize_q_ic_basic, 1., f1, 1) / a; }
	if (baryon_flag > 1) baryon_flag = 0;                                                              // :2072

	// ---- phi (:2180-2214)
	s->plan_source.execute(FFT_FORWARD);
	G.displacement_field(0., pkspline, 0);
	s->plan_phi.execute(FFT_BACKWARD);
	s->phi.updateHalo();                                                                               // phi now finally contains phi

	// ---- B and chi from the particles (:2236-2254)
	projection_init(&s->Bi);
	projection_T0i_project(&s->pcls_cdm, &s->Bi, &s->phi);
	if (baryon_flag) projection_T0i_project(&s->pcls_b, &s->Bi, &s->phi);
	projection_T0i_comm(&s->Bi);
	s->plan_Bi.execute(FFT_FORWARD);
	projectFTvector(s->BiFT, s->BiFT, fourpiG / (double) st.ngrid / (double) st.ngrid);
	s->plan_Bi.execute(FFT_BACKWARD);
	s->Bi.updateHalo();                                                                                // B initialized
	projection_init(&s->Sij);
	projection_Tij_project(&s->pcls_cdm, &s->Sij, a, &s->phi);
	if (baryon_flag) projection_Tij_project(&s->pcls_b, &s->Sij, a, &s->phi);
	projection_Tij_comm(&s->Sij);
	prepareFTsource<Real>(s->phi, s->Sij, s->Sij, 2. * fourpiG / a / (double) st.ngrid / (double) st.ngrid);
	s->plan_Sij.execute(FFT_FORWARD);
	projectFTscalar(s->SijFT, s->scalarFT);
	s->plan_chi.execute(FFT_BACKWARD);
	s->chi.updateHalo();                                                                               // chi now finally contains chi

	// ---- main.cpp:325-333: the maximum velocities feed the first cycle
	s->lat.max(maxvel, 2);
	if (gr > 0) for (int i = 0; i < 2; i++) maxvel[i] /= std::sqrt(maxvel[i] * maxvel[i] + 1.0);
	s->maxvel[0] = maxvel[0]; s->maxvel[1] = baryon_flag ? maxvel[1] : 0.;
	return 0;
}

} // namespace

extern "C" int gevb_ic_cic_kernel(int ngrid, int64_t numpcl, const float * pcldata, int numtile, double * out27)
{
	if (ngrid < 2 || out27 == NULL || numtile < 1) return 1;
	cic_kernel(ngrid, (long) numpcl, pcldata, numtile, out27);
	return 0;
}

extern "C" int gevb_ic_displacement_field(int ngrid, double * potFT, double coeff, int nspline, const double * spline_x, const double * spline_y,
                                          unsigned int seed, int ksphere, int deconvolve_f)
{
	if (ngrid < 4 || (ngrid & 1) || potFT == NULL || nspline < 2 || spline_x == NULL || spline_y == NULL) return 1;
	try
	{
		Spline pk;
		pk.init(spline_x, spline_y, (size_t) nspline);
		realise(ngrid, 0, ngrid, false, potFT, coeff, pk, seed, ksphere, deconvolve_f);
	}
	catch (...) { return 1; }
	return 0;
}

extern "C" int gevb_ic_load_template(const char * filename, int64_t * numpart, float ** pcldata)
{
	if (filename == NULL || numpart == NULL || pcldata == NULL) return 1;
	try
	{
		long n = 0;
		std::vector<float> data;
		if (load_template(filename, n, data) != 0) return 1;
		float * out = (float *) std::malloc(data.size() * sizeof(float));
		if (out == NULL) return 1;
		std::memcpy(out, data.data(), data.size() * sizeof(float));
		*numpart = n; *pcldata = out;
	}
	catch (...) { return 1; }
	return 0;
}

extern "C" int gevb_sim_create_from_settings(gevb_sim ** out, gevb_ctx * ctx, const gevb_settings * st)
{
	if (out == NULL || ctx == NULL || st == NULL) return 1;
	int n = 0;
	gevb_ctx_geometry(ctx, &n, NULL, NULL, NULL, NULL);
	if (n != st->ngrid) { std::fprintf(stderr, " gevb_sim_create_from_settings: the context has Ngrid %d, the settings %d\n", n, st->ngrid); return 1; }
	const double ds[5] = {st->boxsize, st->Cf, st->steplimit, st->z_in, st->z_relax};
	gevb_sim * s = NULL;
	if (gevb_sim_create(&s, ctx, st->gr_flag, st->vector_flag, ds, st->cosmo) != 0) return 1;
	int r = 1;
	try
	{
		// movelimit as the parser leaves it (the file's value, else Ngrid), then the clamp of main.cpp:281-286 inside set_ncdm
		if (gevb_sim_set_ncdm(s, 0, NULL, NULL, NULL, NULL, NULL, 0., st->movelimit) == 0) r = generate(s, *st);
	}
	catch (const gevb_error &) { r = 1; }
	catch (...) { r = 1; }
	if (r != 0) { gevb_sim_destroy(s); return 1; }
	*out = s;
	return 0;
}

extern "C" int gevb_sim_run_settings(gevb_sim * s, const gevb_settings * st, int max_cycles, int * counts3)
{
	if (s == NULL || st == NULL) return 1;
	try
	{
		const std::string pk = std::string(st->output_path) + st->basename_pk, snap = std::string(st->output_path) + st->basename_snapshot;
		const bool gadget = (st->snapshot_mask & 512) != 0;                                            // MASK_GADGET
		return gevb_sim_run(s, st->z_pk, st->pk_mask ? st->num_pk : 0, st->pk_mask, st->numbins, pk.c_str(),
		                    st->z_snapshot, gadget ? st->num_snapshot : 0, st->tracer_factor[0], snap.c_str(), max_cycles, counts3);
	}
	catch (...) { return 1; }
}
