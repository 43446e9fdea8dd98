// =============================================================================
// oracle/latfield2_shim/LATfield2.hpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE
// =============================================================================
// Single-rank, CPU-only stand-in for the parts of LATfield2 v1.1 that the
// reference's hot path consumes.  LATfield2 itself is an
// un-vendored dependency (reference README.md:29) that is absent from
// /root/reference, so its mechanics are RESTATED here from the reference's call
// sites and manual.pdf section 4 ("parity unpinned" at this boundary).  The
// arithmetic of every physics kernel is NOT restated here: oracle/ref_driver.cpp
// #includes the reference's own gevolution.hpp / tools.hpp / background.hpp by
// path against this header, so those formulas are the reference's own text.
//
// Written from scratch for this repository; nothing is copied from LATfield2.
//
// Contract summary (what the reference call sites rely on):
//   Real/Imag           gevolution.hpp:30-32,225,232,317,385; tools.hpp:136,160
//   parallel            main.cpp:152,286,324,462,816 (single rank: identities)
//   Lattice/Site/rKSite main.cpp:213-215; gevolution.hpp:59-70,217,229-240,958-975
//   Field<T>            main.cpp:234-246,518; gevolution.hpp:64,97,434
//   PlanFFT<Cplx>       main.cpp:238-246,477,488 (unnormalised both ways)
//   Particles<...>      gevolution.hpp:935-977; main.cpp:775,798
//   projection_init, *_comm   main.cpp:378,411,435,450
//
// Threading extension (ours, for the CPU baseline only): Site iteration can be
// restricted to a z-range per thread (lf2::set_z_range) so that pointwise
// reference loops may be run slab-parallel under OpenMP.
// =============================================================================
#ifndef LATFIELD2_SHIM_HPP
#define LATFIELD2_SHIM_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstddef>
#include <iostream>
#include <list>
#include <string>
#include <vector>
#include <algorithm>

#ifndef FFT3D
#define FFT3D
#endif

// ---- minimal MPI vocabulary used directly by tools.hpp:169-211 --------------
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_IN_PLACE ((void *) 1)
#define MPI_FLOAT 1
#define MPI_DOUBLE 2
#define MPI_INT 3
#define MPI_SUM 1
static inline int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm) { return 0; }

// ---- MPI-IO vocabulary used by Particles_gevolution.hpp:30-400 (saveGadget2 / loadGadget2), on one rank: plain
//      positioned stdio.  MPI_MODE_CREATE does not truncate, as in MPI.
#include <cstdio>
#include <unistd.h>
#define MPI_UNSIGNED 4
#define MPI_BYTE 5
#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4
#define MPI_INFO_NULL 0
#define MPI_SEEK_SET 0
typedef long long MPI_Offset;
typedef FILE * MPI_File;
struct MPI_Status { int unused; };
static inline size_t lf2_mpi_size(MPI_Datatype t) { return t == MPI_DOUBLE ? 8 : (t == MPI_BYTE ? 1 : 4); }
static inline int MPI_File_open(MPI_Comm, const char * name, int amode, int, MPI_File * fh)
{
	if (amode & MPI_MODE_RDONLY) { *fh = std::fopen(name, "rb"); return *fh ? 0 : 1; }
	*fh = std::fopen(name, "r+b");
	if (*fh == NULL && (amode & MPI_MODE_CREATE)) *fh = std::fopen(name, "w+b");
	return *fh ? 0 : 1;
}
static inline int MPI_File_set_size(MPI_File fh, MPI_Offset size) { std::fflush(fh); return ftruncate(fileno(fh), (off_t) size); }
static inline int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void * buf, int count, MPI_Datatype t, MPI_Status *)
{
	if (fseeko(fh, (off_t) off, SEEK_SET) != 0) return 1;
	return std::fwrite(buf, lf2_mpi_size(t), (size_t) count, fh) == (size_t) count ? 0 : 1;
}
static inline int MPI_File_write_at_all(MPI_File fh, MPI_Offset off, const void * buf, int count, MPI_Datatype t, MPI_Status * st) { return MPI_File_write_at(fh, off, buf, count, t, st); }
static inline int MPI_File_seek(MPI_File fh, MPI_Offset off, int) { return fseeko(fh, (off_t) off, SEEK_SET); }
static inline int MPI_File_read_all(MPI_File fh, void * buf, int count, MPI_Datatype t, MPI_Status *) { return std::fread(buf, lf2_mpi_size(t), (size_t) count, fh) == (size_t) count ? 0 : 1; }
static inline int MPI_File_close(MPI_File * fh) { int r = std::fclose(*fh); *fh = NULL; return r; }

namespace lf2 {
// per-thread restriction of Site iteration to z in [zlo, zhi) (default: all)
struct ZRange { int zlo, zhi; };
inline ZRange & zrange() { static thread_local ZRange r = {0, 1 << 30}; return r; }
inline void set_z_range(int zlo, int zhi) { zrange().zlo = zlo; zrange().zhi = zhi; }
inline void clear_z_range() { zrange().zlo = 0; zrange().zhi = 1 << 30; }
}

namespace LATfield2 {

typedef double Real;

#define FFT_FORWARD 1
#define FFT_BACKWARD (-1)
#define SUM 1
#define MIN 2
#define MAX 3

const int symmetric = 1;
const int unsymmetric = 0;

// ---------------------------------------------------------------------------
// Imag: complex number with the member functions the reference uses
// ---------------------------------------------------------------------------
class Imag
{
public:
	Real re_, im_;
	Imag() : re_(0.), im_(0.) {}
	Imag(Real r, Real i) : re_(r), im_(i) {}
	Real & real() { return re_; }
	Real & imag() { return im_; }
	const Real & real() const { return re_; }
	const Real & imag() const { return im_; }
	Imag conj() const { return Imag(re_, -im_); }
	Real norm() const { return re_ * re_ + im_ * im_; }
	Imag operator-() const { return Imag(-re_, -im_); }
	Imag operator+(const Imag & b) const { return Imag(re_ + b.re_, im_ + b.im_); }
	Imag operator-(const Imag & b) const { return Imag(re_ - b.re_, im_ - b.im_); }
	Imag operator*(const Imag & b) const { return Imag(re_ * b.re_ - im_ * b.im_, re_ * b.im_ + im_ * b.re_); }
	Imag operator/(const Imag & b) const { Real d = b.re_ * b.re_ + b.im_ * b.im_; return Imag((re_ * b.re_ + im_ * b.im_) / d, (im_ * b.re_ - re_ * b.im_) / d); }
	Imag operator+(Real b) const { return Imag(re_ + b, im_); }
	Imag operator-(Real b) const { return Imag(re_ - b, im_); }
	Imag operator*(Real b) const { return Imag(re_ * b, im_ * b); }
	Imag operator/(Real b) const { return Imag(re_ / b, im_ / b); }
	Imag & operator+=(const Imag & b) { re_ += b.re_; im_ += b.im_; return *this; }
	Imag & operator-=(const Imag & b) { re_ -= b.re_; im_ -= b.im_; return *this; }
	Imag & operator*=(const Imag & b) { *this = *this * b; return *this; }
	Imag & operator/=(const Imag & b) { *this = *this / b; return *this; }
	Imag & operator+=(Real b) { re_ += b; return *this; }
	Imag & operator-=(Real b) { re_ -= b; return *this; }
	Imag & operator*=(Real b) { re_ *= b; im_ *= b; return *this; }
	Imag & operator/=(Real b) { re_ /= b; im_ /= b; return *this; }
};
inline Imag operator*(Real a, const Imag & b) { return Imag(a * b.re_, a * b.im_); }
inline Imag operator+(Real a, const Imag & b) { return Imag(a + b.re_, b.im_); }
inline Imag operator-(Real a, const Imag & b) { return Imag(a - b.re_, -b.im_); }
inline Imag operator/(Real a, const Imag & b) { return Imag(a, 0.) / b; }

// ---------------------------------------------------------------------------
// parallel: single-rank stand-in for Parallel2d
// ---------------------------------------------------------------------------
class Parallel2d
{
	int grid_rank_[2];
	int grid_size_[2];
	MPI_Comm comms_[1] = {0};
	std::vector<char> mailbox_;
public:
	Parallel2d() { grid_rank_[0] = grid_rank_[1] = 0; grid_size_[0] = grid_size_[1] = 1; }
	void initialize(int, int) {}
	int rank() const { return 0; }
	int size() const { return 1; }
	bool isRoot() const { return true; }
	int root() const { return 0; }
	int * grid_rank() { return grid_rank_; }
	int * grid_size() { return grid_size_; }
	MPI_Comm lat_world_comm() const { return 0; }
	MPI_Comm * dim0_comm() { return comms_; }
	MPI_Comm * dim1_comm() { return comms_; }
	// point-to-point on one rank: the ring of size one that saveGadget2 builds (Particles_gevolution.hpp:86-116)
	// sends to "the next rank" and receives from "the previous one" -- both are this rank, so a mailbox does it
	template <class T> void send(T & v, int) { mailbox_.assign((const char *) &v, (const char *) &v + sizeof(T)); }
	template <class T> void receive(T & v, int) { if (mailbox_.size() == sizeof(T)) std::memcpy(&v, mailbox_.data(), sizeof(T)); }
	template <class T> void send_dim0(T & v, int to) { send(v, to); }
	template <class T> void receive_dim0(T & v, int from) { receive(v, from); }
	void abortForce() { std::cerr << "parallel.abortForce()" << std::endl; std::exit(-1); }
	void barrier() {}
	template <class T> void sum(T &) {}
	template <class T> void sum(T *, int) {}
	template <class T> void max(T &) {}
	template <class T> void max(T *, int) {}
	template <class T> void min(T &) {}
	template <class T> void min(T *, int) {}
	template <class T> void broadcast(T &, int) {}
	template <class T> void broadcast(T *, int, int) {}
	template <class T> void broadcast_dim0(T &, int) {}
	template <class T> void broadcast_dim0(T *, int, int) {}
	template <class T> void broadcast_dim1(T &, int) {}
	template <class T> void broadcast_dim1(T *, int, int) {}
};

// a real object (not a macro: OpenMP pragmas are macro-expanded); needs -std=c++17
inline Parallel2d parallel;
#define COUT if (LATfield2::parallel.isRoot()) std::cout

// ---------------------------------------------------------------------------
// Lattice: 3-D periodic lattice with a halo of width `halo` in every dimension
// memory order: x (dim 0) fastest, then y, then z; halo cells included
// ---------------------------------------------------------------------------
class Lattice
{
	int dim_;
	int size_[3];
	int halo_;
	long jump_[3];
	long sitesLocal_, sitesLocalGross_, siteFirst_, siteLast_;
	int coordSkip_[2];
public:
	Lattice() : dim_(0), halo_(0) { size_[0] = size_[1] = size_[2] = 0; coordSkip_[0] = coordSkip_[1] = 0; }
	Lattice(int dim, const int * size, int halo) { initialize(dim, size, halo); }
	Lattice(int dim, const int size, int halo) { int s[3] = {size, size, size}; initialize(dim, s, halo); }
	void initialize(int dim, const int * size, int halo)
	{
		if (dim != 3) { std::cerr << "LATfield2 shim: only dim=3 is supported" << std::endl; std::exit(-1); }
		dim_ = dim; halo_ = halo;
		for (int i = 0; i < 3; i++) size_[i] = size[i];
		jump_[0] = 1;
		jump_[1] = size_[0] + 2 * halo_;
		jump_[2] = jump_[1] * (size_[1] + 2 * halo_);
		sitesLocal_ = (long) size_[0] * size_[1] * size_[2];
		sitesLocalGross_ = jump_[2] * (size_[2] + 2 * halo_);
		siteFirst_ = halo_ * (jump_[0] + jump_[1] + jump_[2]);
		siteLast_ = sitesLocalGross_ - 1 - siteFirst_;
		coordSkip_[0] = coordSkip_[1] = 0;
	}
	void initialize(int dim, const int size, int halo) { int s[3] = {size, size, size}; initialize(dim, s, halo); }
	// Fourier-space companion of a real lattice: (N/2+1, N, N), manual.pdf section 4
	void initializeRealFFT(Lattice & real, int halo)
	{
		int s[3] = {real.size(0) / 2 + 1, real.size(1), real.size(2)};
		initialize(3, s, halo);
	}
	int dim() const { return dim_; }
	int halo() const { return halo_; }
	int size(int i) const { return size_[i]; }
	const int * size() const { return size_; }
	int sizeLocal(int i) const { return size_[i]; }
	long jump(int i) const { return jump_[i]; }
	long sites() const { return sitesLocal_; }
	long sitesLocal() const { return sitesLocal_; }
	long sitesLocalGross() const { return sitesLocalGross_; }
	long siteFirst() const { return siteFirst_; }
	long siteLast() const { return siteLast_; }
	int * coordSkip() { return coordSkip_; }
	long indexOf(int x, int y, int z) const { return (x + halo_) * jump_[0] + (y + halo_) * jump_[1] + (z + halo_) * jump_[2]; }
};

// ---------------------------------------------------------------------------
// Site / rKSite: cursor over the bulk sites of a lattice (x fastest)
// ---------------------------------------------------------------------------
class Site
{
protected:
	Lattice * lattice_;
	long index_;
	int c_[3];       // cached bulk coordinates of the cursor (valid while iterating)
	int zend_;
public:
	Site() : lattice_(NULL), index_(0), zend_(0) { c_[0] = c_[1] = c_[2] = 0; }
	Site(Lattice & lat) { initialize(lat); }
	Site(Lattice & lat, long index) { initialize(lat, index); }
	void initialize(Lattice & lat) { lattice_ = &lat; index_ = lat.siteFirst(); c_[0] = c_[1] = c_[2] = 0; zend_ = lat.size(2); }
	void initialize(Lattice & lat, long index) { lattice_ = &lat; zend_ = lat.size(2); setIndex(index); }
	void first()
	{
		const lf2::ZRange & r = lf2::zrange();
		int z0 = std::min(r.zlo, lattice_->size(2));
		zend_ = std::min(r.zhi, lattice_->size(2));
		c_[0] = 0; c_[1] = 0; c_[2] = z0;
		index_ = lattice_->indexOf(0, 0, z0);
	}
	bool test() const { return c_[2] < zend_; }
	void next()
	{
		index_++;
		if (++c_[0] == lattice_->size(0))
		{
			c_[0] = 0;
			index_ += 2 * lattice_->halo();
			if (++c_[1] == lattice_->size(1))
			{
				c_[1] = 0;
				index_ += 2 * lattice_->halo() * lattice_->jump(1);
				++c_[2];
			}
		}
	}
	long index() const { return index_; }
	void setIndex(long idx)
	{
		index_ = idx;
		long r = idx;
		int h = lattice_->halo();
		c_[2] = (int) (r / lattice_->jump(2)) - h; r %= lattice_->jump(2);
		c_[1] = (int) (r / lattice_->jump(1)) - h; r %= lattice_->jump(1);
		c_[0] = (int) r - h;
	}
	int coord(int i) const { return c_[i]; }
	int coordLocal(int i) const { return c_[i]; }
	bool setCoord(int x, int y, int z) { c_[0] = x; c_[1] = y; c_[2] = z; index_ = lattice_->indexOf(x, y, z); return true; }
	bool setCoord(const int * r) { return setCoord(r[0], r[1], r[2]); }
	bool setCoordLocal(const int * r) { return setCoord(r[0], r[1], r[2]); }
	Lattice & lattice() const { return *lattice_; }
	// neighbour arithmetic: may land in the halo, never wraps (periodicity lives in the halo)
	Site operator+(int dir) const { Site s(*this); s.index_ += lattice_->jump(dir); s.c_[dir]++; return s; }
	Site operator-(int dir) const { Site s(*this); s.index_ -= lattice_->jump(dir); s.c_[dir]--; return s; }
};

class rKSite : public Site
{
public:
	rKSite() : Site() {}
	rKSite(Lattice & lat) : Site(lat) {}
	rKSite(Lattice & lat, long index) : Site(lat, index) {}
	rKSite operator+(int dir) const { rKSite s(*this); s.index_ += lattice_->jump(dir); s.c_[dir]++; return s; }
	rKSite operator-(int dir) const { rKSite s(*this); s.index_ -= lattice_->jump(dir); s.c_[dir]--; return s; }
};

// ---------------------------------------------------------------------------
// Field<T>: components contiguous per site (array of structs), halo included
// ---------------------------------------------------------------------------
template <class FieldType>
class Field
{
	Lattice * lattice_;
	FieldType * data_;
	int components_, rows_, cols_, symmetry_;
	void release() { if (data_ != NULL) { delete[] data_; data_ = NULL; } }
public:
	Field() : lattice_(NULL), data_(NULL), components_(0), rows_(0), cols_(0), symmetry_(unsymmetric) {}
	Field(Lattice & lat, int comps = 1) : lattice_(NULL), data_(NULL) { initialize(lat, comps); alloc(); }
	~Field() { release(); }
	void initialize(Lattice & lat, int comps = 1)
	{
		release();
		lattice_ = &lat; components_ = comps; rows_ = comps; cols_ = 1; symmetry_ = unsymmetric;
	}
	void initialize(Lattice & lat, int rows, int cols, int sym)
	{
		release();
		lattice_ = &lat; rows_ = rows; cols_ = cols; symmetry_ = sym;
		components_ = (sym == symmetric) ? (rows * (rows + 1)) / 2 : rows * cols;
	}
	void alloc()
	{
		if (data_ == NULL) data_ = new FieldType[(size_t) lattice_->sitesLocalGross() * components_]();
	}
	void dealloc() { release(); }
	FieldType * data() { return data_; }
	Lattice & lattice() const { return *lattice_; }
	int components() const { return components_; }
	int rows() const { return rows_; }
	int cols() const { return cols_; }
	int symmetry() const { return symmetry_; }
	int compIndex(int i, int j) const
	{
		if (symmetry_ == symmetric)
		{
			int lo = i < j ? i : j, hi = i < j ? j : i;
			return (hi - lo) + (lo * (2 * rows_ + 1 - lo)) / 2;   // (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
		}
		return i * cols_ + j;
	}
	FieldType & operator()(long index) { return data_[index * components_]; }
	FieldType & operator()(long index, int c) { return data_[index * components_ + c]; }
	FieldType & operator()(const Site & s) { return data_[s.index() * components_]; }
	FieldType & operator()(const Site & s, int c) { return data_[s.index() * components_ + c]; }
	FieldType & operator()(const Site & s, int i, int j) { return data_[s.index() * components_ + compIndex(i, j)]; }
	// periodic ghost fill of width halo in all three dimensions
	void updateHalo()
	{
		const int h = lattice_->halo();
		if (h == 0) return;
		const int nx = lattice_->size(0), ny = lattice_->size(1), nz = lattice_->size(2);
		for (int z = -h; z < nz + h; z++)
			for (int y = -h; y < ny + h; y++)
				for (int x = -h; x < nx + h; x++)
				{
					if (x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz) continue;
					int xs = ((x % nx) + nx) % nx, ys = ((y % ny) + ny) % ny, zs = ((z % nz) + nz) % nz;
					long dst = lattice_->indexOf(x, y, z), src = lattice_->indexOf(xs, ys, zs);
					for (int c = 0; c < components_; c++) data_[dst * components_ + c] = data_[src * components_ + c];
				}
	}
};

// ---------------------------------------------------------------------------
// 1-D complex FFT kernels used by PlanFFT (own code; radix-2 + naive fallback)
// ---------------------------------------------------------------------------
namespace fftimpl {
struct Plan1d
{
	int n; bool pow2;
	std::vector<double> wr, wi;      // e^{-2 pi i k / n}
	std::vector<int> rev;
	explicit Plan1d(int n_) : n(n_)
	{
		pow2 = (n > 0) && ((n & (n - 1)) == 0);
		wr.resize(n); wi.resize(n);
		for (int k = 0; k < n; k++) { double a = -2.0 * M_PI * (double) k / (double) n; wr[k] = cos(a); wi[k] = sin(a); }
		if (pow2)
		{
			rev.resize(n);
			int bits = 0; while ((1 << bits) < n) bits++;
			for (int i = 0; i < n; i++) { int r = 0; for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b); rev[i] = r; }
		}
	}
	// in-place transform of (re,im) with stride 1; sign=-1 forward (e^{-ikx}), +1 backward
	void run(double * re, double * im, int sign, double * tr, double * ti) const
	{
		if (pow2)
		{
			for (int i = 0; i < n; i++) { int r = rev[i]; if (r > i) { std::swap(re[i], re[r]); std::swap(im[i], im[r]); } }
			for (int len = 2; len <= n; len <<= 1)
			{
				int half = len >> 1, step = n / len;
				for (int s = 0; s < n; s += len)
					for (int k = 0; k < half; k++)
					{
						double c = wr[k * step], d = (sign < 0) ? wi[k * step] : -wi[k * step];
						double xr = re[s + k + half] * c - im[s + k + half] * d;
						double xi = re[s + k + half] * d + im[s + k + half] * c;
						re[s + k + half] = re[s + k] - xr; im[s + k + half] = im[s + k] - xi;
						re[s + k] += xr; im[s + k] += xi;
					}
			}
		}
		else
		{
			for (int k = 0; k < n; k++)
			{
				double sr = 0., si = 0.;
				for (int j = 0; j < n; j++)
				{
					int idx = (int) (((long) j * k) % n);
					double c = wr[idx], d = (sign < 0) ? wi[idx] : -wi[idx];
					sr += re[j] * c - im[j] * d; si += re[j] * d + im[j] * c;
				}
				tr[k] = sr; ti[k] = si;
			}
			for (int k = 0; k < n; k++) { re[k] = tr[k]; im[k] = ti[k]; }
		}
	}
};
}

// ---------------------------------------------------------------------------
// PlanFFT<Imag>: per-component 3-D r2c / c2r, unnormalised both ways
//   forward : F(k) = sum_x f(x) e^{-2 pi i k.x/N}, kx in [0,N/2]
//   backward: c2r of the half spectrum (Hermitian completion along x; the
//             imaginary parts of self-conjugate modes are ignored, like FFTW)
// The halo of the real field is neither read nor refreshed.
// ---------------------------------------------------------------------------
template <class CplxType>
class PlanFFT
{
	Field<Real> * rfield_;
	Field<CplxType> * kfield_;
	int n_[3];
public:
	PlanFFT() : rfield_(NULL), kfield_(NULL) {}
	PlanFFT(Field<Real> * r, Field<CplxType> * k) { initialize(r, k); }
	void initialize(Field<Real> * r, Field<CplxType> * k)
	{
		rfield_ = r; kfield_ = k;
		r->alloc(); k->alloc();
		for (int i = 0; i < 3; i++) n_[i] = r->lattice().size(i);
	}
	void execute(int dir)
	{
		const int nx = n_[0], ny = n_[1], nz = n_[2], nxh = nx / 2 + 1;
		const int comps = rfield_->components();
		Lattice & rl = rfield_->lattice();
		Lattice & kl = kfield_->lattice();
		fftimpl::Plan1d px(nx), py(ny), pz(nz);
		const int nmax = std::max(nx, std::max(ny, nz));
		std::vector<double> wr((size_t) nxh * ny * nz), wi((size_t) nxh * ny * nz);
		for (int c = 0; c < comps; c++)
		{
			if (dir == FFT_FORWARD)
			{
				#pragma omp parallel
				{
					std::vector<double> re(nmax), im(nmax), tr(nmax), ti(nmax);
					#pragma omp for collapse(2)
					for (int z = 0; z < nz; z++) for (int y = 0; y < ny; y++)
					{
						for (int x = 0; x < nx; x++) { re[x] = (*rfield_)(rl.indexOf(x, y, z), c); im[x] = 0.; }
						px.run(re.data(), im.data(), -1, tr.data(), ti.data());
						size_t o = ((size_t) z * ny + y) * nxh;
						for (int x = 0; x < nxh; x++) { wr[o + x] = re[x]; wi[o + x] = im[x]; }
					}
					#pragma omp for collapse(2)
					for (int z = 0; z < nz; z++) for (int x = 0; x < nxh; x++)
					{
						for (int y = 0; y < ny; y++) { size_t o = ((size_t) z * ny + y) * nxh + x; re[y] = wr[o]; im[y] = wi[o]; }
						py.run(re.data(), im.data(), -1, tr.data(), ti.data());
						for (int y = 0; y < ny; y++) { size_t o = ((size_t) z * ny + y) * nxh + x; wr[o] = re[y]; wi[o] = im[y]; }
					}
					#pragma omp for collapse(2)
					for (int y = 0; y < ny; y++) for (int x = 0; x < nxh; x++)
					{
						for (int z = 0; z < nz; z++) { size_t o = ((size_t) z * ny + y) * nxh + x; re[z] = wr[o]; im[z] = wi[o]; }
						pz.run(re.data(), im.data(), -1, tr.data(), ti.data());
						for (int z = 0; z < nz; z++) (*kfield_)(kl.indexOf(x, y, z), c) = CplxType(re[z], im[z]);
					}
				}
			}
			else
			{
				#pragma omp parallel
				{
					std::vector<double> re(nmax), im(nmax), tr(nmax), ti(nmax);
					#pragma omp for collapse(2)
					for (int y = 0; y < ny; y++) for (int x = 0; x < nxh; x++)
					{
						for (int z = 0; z < nz; z++) { const CplxType & v = (*kfield_)(kl.indexOf(x, y, z), c); re[z] = v.real(); im[z] = v.imag(); }
						pz.run(re.data(), im.data(), +1, tr.data(), ti.data());
						for (int z = 0; z < nz; z++) { size_t o = ((size_t) z * ny + y) * nxh + x; wr[o] = re[z]; wi[o] = im[z]; }
					}
					#pragma omp for collapse(2)
					for (int z = 0; z < nz; z++) for (int x = 0; x < nxh; x++)
					{
						for (int y = 0; y < ny; y++) { size_t o = ((size_t) z * ny + y) * nxh + x; re[y] = wr[o]; im[y] = wi[o]; }
						py.run(re.data(), im.data(), +1, tr.data(), ti.data());
						for (int y = 0; y < ny; y++) { size_t o = ((size_t) z * ny + y) * nxh + x; wr[o] = re[y]; wi[o] = im[y]; }
					}
					#pragma omp for collapse(2)
					for (int z = 0; z < nz; z++) for (int y = 0; y < ny; y++)
					{
						size_t o = ((size_t) z * ny + y) * nxh;
						re[0] = wr[o]; im[0] = 0.;
						for (int x = 1; x < nxh; x++) { re[x] = wr[o + x]; im[x] = wi[o + x]; }
						if (nx % 2 == 0) im[nx / 2] = 0.;
						for (int x = nxh; x < nx; x++) { re[x] = re[nx - x]; im[x] = -im[nx - x]; }
						px.run(re.data(), im.data(), +1, tr.data(), ti.data());
						for (int x = 0; x < nx; x++) (*rfield_)(rl.indexOf(x, y, z), c) = re[x];
					}
				}
			}
		}
	}
};

// ---------------------------------------------------------------------------
// Particles
// ---------------------------------------------------------------------------
struct part_simple
{
	long ID;
	Real pos[3];
	Real vel[3];
};

struct part_simple_info
{
	double mass;
	int relativistic;
	char type_name[64];
};

struct part_simple_dataType { };

template <typename part>
struct partList
{
	int size;
	std::list<part> parts;
	partList() : size(0) {}
};

template <typename part, typename part_info, typename part_dataType>
class Particles
{
protected:
	Lattice lat_part_;
	Field<partList<part> > field_part_;
	part_info part_global_info_;
	Real lat_resolution_;
	Real boxSize_[3];
	long numParticles_;

	// cell of a (wrapped) position: floor(pos/dx), clamped to the lattice
	int cellOf(Real p, int n) const
	{
		int c = (int) std::floor(p / lat_resolution_);
		if (c >= n) c = n - 1;
		if (c < 0) c = 0;
		return c;
	}
public:
	Particles() : lat_resolution_(0.), numParticles_(0) {}
	void initialize(part_info info, part_dataType, Lattice * lat, Real boxSize[3])
	{
		part_global_info_ = info;
		int s[3] = {lat->size(0), lat->size(1), lat->size(2)};
		lat_part_.initialize(3, s, 0);
		field_part_.initialize(lat_part_, 1);
		field_part_.alloc();
		for (int i = 0; i < 3; i++) boxSize_[i] = boxSize[i];
		lat_resolution_ = boxSize[0] / (Real) lat->size(0);
		numParticles_ = 0;
	}
	Lattice & lattice() { return lat_part_; }
	Field<partList<part> > & field() { return field_part_; }
	Real res() const { return lat_resolution_; }
	part_info * parts_info() { return &part_global_info_; }
	size_t mass_offset() const { return offsetof(part_info, mass); }
	long numParticles() const { return numParticles_; }

	// periodic wrap convention ("unverified edge semantics" (1)):
	//   p' = p - floor(p/L)*L ; a result that rounds to L is mapped to 0
	static Real wrapPos(Real p, Real L)
	{
		Real w = p - std::floor(p / L) * L;
		if (w >= L) w = 0.;
		return w;
	}

	bool addParticle_global(part p)
	{
		int c[3];
		for (int l = 0; l < 3; l++) c[l] = cellOf(p.pos[l], lat_part_.size(l));
		partList<part> & cell = field_part_(lat_part_.indexOf(c[0], c[1], c[2]));
		cell.parts.push_back(p);
		cell.size++;
		numParticles_++;
		return true;
	}

	typedef Real (*vel_fn)(double, double, part *, double *, part_info, Field<Real> **, Site *, int, double *, double *, int);
	typedef void (*pos_fn)(double, double, part *, double *, part_info, Field<Real> **, Site *, int, double *, double *, int);

private:
	static void reduceInto(double * acc, const double * val, const int * reduce_type, int n)
	{
		for (int i = 0; i < n; i++)
		{
			int t = (reduce_type != NULL) ? reduce_type[i] : SUM;
			if (t == SUM) acc[i] += val[i];
			else if (t == MIN) { if (val[i] < acc[i]) acc[i] = val[i]; }
			else if (t == MAX) { if (val[i] > acc[i]) acc[i] = val[i]; }
		}
	}
public:
	// kick driver (main.cpp:775): per particle ref_dist = frac(pos/dx); returns sqrt(max fn)
	Real updateVel(vel_fn fn, double dtau, Field<Real> ** fields, int nfields, double * params = NULL, double * output = NULL, int * reduce_type = NULL, int noutput = 0)
	{
		Site xPart(lat_part_);
		std::vector<Site> sites(nfields > 0 ? nfields : 1);
		for (int i = 0; i < nfields; i++) sites[i].initialize(fields[i]->lattice());
		Real maxv2 = 0.;
		double frac[3], ipart;
		typename std::list<part>::iterator it;
		for (xPart.first(); xPart.test(); xPart.next())
		{
			partList<part> & cell = field_part_(xPart);
			if (cell.size == 0) continue;
			for (int i = 0; i < nfields; i++) sites[i].setCoord(xPart.coord(0), xPart.coord(1), xPart.coord(2));
			for (it = cell.parts.begin(); it != cell.parts.end(); ++it)
			{
				for (int l = 0; l < 3; l++) frac[l] = modf((*it).pos[l] / lat_resolution_, &ipart);
				Real v2 = fn(dtau, lat_resolution_, &(*it), frac, part_global_info_, fields, sites.data(), nfields, params, output, noutput);
				if (v2 > maxv2) maxv2 = v2;
			}
		}
		(void) reduce_type;
		return sqrt(maxv2);
	}

	// drift driver (main.cpp:798): callback, periodic wrap, re-file under floor(pos/dx)
	void moveParticles(pos_fn fn, double dtau, Field<Real> ** fields = NULL, int nfields = 0, double * params = NULL, double * output = NULL, int * reduce_type = NULL, int noutput = 0)
	{
		Site xPart(lat_part_);
		std::vector<Site> sites(nfields > 0 ? nfields : 1);
		for (int i = 0; i < nfields; i++) sites[i].initialize(fields[i]->lattice());
		double frac[3], ipart;
		std::vector<double> tmp(noutput > 0 ? noutput : 1);
		std::list<part> moved;                       // particles that left their cell
		typename std::list<part>::iterator it, cur;
		for (xPart.first(); xPart.test(); xPart.next())
		{
			partList<part> & cell = field_part_(xPart);
			if (cell.size == 0) continue;
			for (int i = 0; i < nfields; i++) sites[i].setCoord(xPart.coord(0), xPart.coord(1), xPart.coord(2));
			for (it = cell.parts.begin(); it != cell.parts.end(); )
			{
				cur = it++;
				for (int l = 0; l < 3; l++) frac[l] = modf((*cur).pos[l] / lat_resolution_, &ipart);
				if (noutput > 0)
				{
					for (int i = 0; i < noutput; i++) tmp[i] = output[i];
					fn(dtau, lat_resolution_, &(*cur), frac, part_global_info_, fields, sites.data(), nfields, params, tmp.data(), noutput);
					reduceInto(output, tmp.data(), reduce_type, noutput);
				}
				else
					fn(dtau, lat_resolution_, &(*cur), frac, part_global_info_, fields, sites.data(), nfields, params, output, noutput);
				bool same = true;
				for (int l = 0; l < 3; l++)
				{
					(*cur).pos[l] = wrapPos((*cur).pos[l], boxSize_[l]);
					if (cellOf((*cur).pos[l], lat_part_.size(l)) != xPart.coord(l)) same = false;
				}
				if (!same)
				{
					moved.splice(moved.end(), cell.parts, cur);
					cell.size--;
				}
			}
		}
		for (it = moved.begin(); it != moved.end(); )
		{
			cur = it++;
			int c[3];
			for (int l = 0; l < 3; l++) c[l] = cellOf((*cur).pos[l], lat_part_.size(l));
			partList<part> & cell = field_part_(lat_part_.indexOf(c[0], c[1], c[2]));
			cell.parts.splice(cell.parts.end(), moved, cur);
			cell.size++;
		}
	}
};

// ---------------------------------------------------------------------------
// projection helpers (main.cpp:378,402,411,435,450)
// ---------------------------------------------------------------------------
template <class FieldType>
void projection_init(Field<FieldType> * f)
{
	size_t n = (size_t) f->lattice().sitesLocalGross() * f->components();
	FieldType * d = f->data();
	for (size_t i = 0; i < n; i++) d[i] = FieldType(0);
}

// fold every upper-halo layer into the periodically corresponding first bulk layer
template <class FieldType>
void projection_fold_upper_halo(Field<FieldType> * f)
{
	Lattice & lat = f->lattice();
	const int h = lat.halo();
	const int nx = lat.size(0), ny = lat.size(1), nz = lat.size(2);
	const int comps = f->components();
	FieldType * d = f->data();
	for (int z = 0; z < nz + h; z++)
		for (int y = 0; y < ny + h; y++)
			for (int x = 0; x < nx + h; x++)
			{
				if (x < nx && y < ny && z < nz) continue;
				long src = lat.indexOf(x, y, z), dst = lat.indexOf(x % nx, y % ny, z % nz);
				for (int c = 0; c < comps; c++) { d[dst * comps + c] += d[src * comps + c]; d[src * comps + c] = FieldType(0); }
			}
}
template <class FieldType> void scalarProjectionCIC_comm(Field<FieldType> * f) { projection_fold_upper_halo(f); }
template <class FieldType> void vectorProjectionCICNGP_comm(Field<FieldType> * f) { projection_fold_upper_halo(f); }
template <class FieldType> void symtensorProjectionCICNGP_comm(Field<FieldType> * f) { projection_fold_upper_halo(f); }

// plain (Newtonian) CIC mass deposit: corner += w * mass / dx^3  (main.cpp:402)
template <typename part, typename part_info, typename part_dataType>
void scalarProjectionCIC_project(Particles<part, part_info, part_dataType> * pcls, Field<Real> * rho)
{
	Site xPart(pcls->lattice());
	Site xField(rho->lattice());
	const Real dx = pcls->res();
	const double mass = *(double *) ((char *) pcls->parts_info() + pcls->mass_offset()) / (dx * dx * dx);
	typename std::list<part>::iterator it;
	for (xPart.first(), xField.first(); xPart.test(); xPart.next(), xField.next())
	{
		partList<part> & cell = pcls->field()(xPart);
		if (cell.size == 0) continue;
		Real cube[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (it = cell.parts.begin(); it != cell.parts.end(); ++it)
		{
			Real up[3], dn[3];
			for (int i = 0; i < 3; i++) { up[i] = ((*it).pos[i] - xPart.coord(i) * dx) / dx; dn[i] = 1. - up[i]; }
			cube[0] += dn[0] * dn[1] * dn[2]; cube[1] += dn[0] * dn[1] * up[2];
			cube[2] += dn[0] * up[1] * dn[2]; cube[3] += dn[0] * up[1] * up[2];
			cube[4] += up[0] * dn[1] * dn[2]; cube[5] += up[0] * dn[1] * up[2];
			cube[6] += up[0] * up[1] * dn[2]; cube[7] += up[0] * up[1] * up[2];
		}
		(*rho)(xField) += cube[0] * mass;       (*rho)(xField + 2) += cube[1] * mass;
		(*rho)(xField + 1) += cube[2] * mass;   (*rho)(xField + 1 + 2) += cube[3] * mass;
		(*rho)(xField + 0) += cube[4] * mass;   (*rho)(xField + 0 + 2) += cube[5] * mass;
		(*rho)(xField + 0 + 1) += cube[6] * mass; (*rho)(xField + 0 + 1 + 2) += cube[7] * mass;
	}
}

} // namespace LATfield2

#endif
