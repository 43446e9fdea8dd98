// settings.cpp -- the reference's settings.ini, as far as the hot path, its outputs and the basic IC generator need it
//
// Restates the subset of parser.hpp that a run of main.cpp with "IC generator = basic" consults: the line format of
// readline (parser.hpp:40-105: "name = value  # comment"), first-match lookup of parseParameter (:305-330) and the keys,
// defaults and derived quantities of parseMetadata (:759-1800).  The same file therefore drives the reference and this
// library.  Everything outside that subset (mPk file, CLASS, lightcones, ncdm species from the generator, restart from
// disk) is reported as unsupported instead of being half-read.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gevb.h"
#include "background.hpp"

namespace {

struct Entry { std::string name, value; };

// one line of the file: true if it declares a parameter (parser.hpp:58-105)
bool split_line(const std::string & line, Entry & e)
{
	const size_t eq = line.find('=');
	if (eq == std::string::npos || eq == 0) return false;
	const size_t hash = line.find('#');
	if (hash != std::string::npos && hash < eq) return false;
	auto blank = [](char c) { return c == ' ' || c == '\t'; };
	size_t l = 0, r = eq;
	while (l < eq && blank(line[l])) l++;
	while (r > l && blank(line[r - 1])) r--;
	if (r <= l) return false;
	e.name = line.substr(l, r - l);
	l = eq + 1;
	r = hash == std::string::npos ? line.size() : hash;
	while (l < r && blank(line[l])) l++;
	while (r > l && (blank(line[r - 1]) || line[r - 1] == '\n' || line[r - 1] == '\r')) r--;
	if (r <= l) return false;
	e.value = line.substr(l, r - l);
	return true;
}

struct Params
{
	std::vector<Entry> list;
	const char * find(const char * name) const
	{
		for (const Entry & e : list) if (e.name == name) return e.value.c_str();
		return NULL;
	}
	bool get(const char * name, double & v) const { const char * s = find(name); return s != NULL && std::sscanf(s, "%lf", &v) == 1; }
	bool get(const char * name, int & v) const { const char * s = find(name); return s != NULL && std::sscanf(s, "%d", &v) == 1; }
	bool get(const char * name, std::string & v) const { const char * s = find(name); if (s == NULL) return false; v = s; return true; }
	// comma-separated list (parser.hpp: parseParameter for arrays)
	std::vector<std::string> items(const char * name) const
	{
		std::vector<std::string> out;
		const char * s = find(name);
		if (s == NULL) return out;
		std::string cur;
		for (const char * q = s; ; q++)
		{
			if (*q == ',' || *q == 0)
			{
				size_t b = cur.find_first_not_of(" \t"), e = cur.find_last_not_of(" \t");
				if (b != std::string::npos) out.push_back(cur.substr(b, e - b + 1));
				cur.clear();
				if (*q == 0) break;
			}
			else cur += *q;
		}
		return out;
	}
};

// parseFieldSpecifiers (parser.hpp:620-735): output names -> MASK_* bits (metadata.hpp:56-70)
int field_mask(const Params & P, const char * name)
{
	static const struct { const char * item; int bit; } names[] = {
		{"Phi", 1}, {"phi", 1}, {"Chi", 2}, {"chi", 2}, {"Pot", 4}, {"pot", 4}, {"Psi_N", 4}, {"psi_N", 4}, {"PsiN", 4}, {"psiN", 4},
		{"B", 8}, {"Bi", 8}, {"P", 256}, {"p", 256}, {"T00", 16}, {"rho", 16}, {"Tij", 32}, {"rho_N", 64}, {"rhoN", 64}, {"hij", 128}, {"GW", 128},
		{"Gadget", 512}, {"Gadget2", 512}, {"gadget", 512}, {"gadget2", 512}, {"multi-Gadget", 512 | 16384}, {"multi-Gadget2", 512 | 16384},
		{"multi-gadget", 512 | 16384}, {"multi-gadget2", 512 | 16384}, {"Particles", 1024}, {"particles", 1024}, {"pcls", 1024}, {"part", 1024},
		{"cross", 2048}, {"X-spectra", 2048}, {"x-spectra", 2048}, {"delta", 4096}, {"Ds", 4096}, {"D_s", 4096}, {"delta_N", 8192}, {"deltaN", 8192}};
	int mask = 0;
	for (const std::string & it : P.items(name))
		for (const auto & n : names) if (it == n.item) mask |= n.bit;
	return mask;
}

int fail(const char * what) { std::fprintf(stderr, " gevb_settings_read: %s\n", what); return 1; }

void copy_str(char * dst, size_t cap, const std::string & s) { std::snprintf(dst, cap, "%s", s.c_str()); }

} // namespace

extern "C" int gevb_settings_read(const char * filename, const char * overrides, gevb_settings * st)
{
	if (filename == NULL || st == NULL) return 1;
	std::memset(st, 0, sizeof(*st));
	Params P;
	try
	{
		FILE * f = std::fopen(filename, "r");
		if (f == NULL) { std::fprintf(stderr, " gevb_settings_read: unable to open parameter file %s\n", filename); return 1; }
		char line[2048];
		while (std::fgets(line, sizeof(line), f) != NULL) { Entry e; if (split_line(line, e)) P.list.push_back(e); }
		std::fclose(f);
		if (P.list.empty()) return fail("no valid data found in the parameter file");
		// overrides take the place of the file's line of the same name (first match wins in the lookup, so they go first)
		if (overrides != NULL)
		{
			std::vector<Entry> first;
			std::string cur;
			for (const char * q = overrides; ; q++)
			{
				if (*q == '\n' || *q == 0) { Entry e; if (split_line(cur, e)) first.push_back(e); cur.clear(); if (*q == 0) break; }
				else cur += *q;
			}
			P.list.insert(P.list.begin(), first.begin(), first.end());
		}
	}
	catch (...) { return 1; }
	std::string s;
	double tmp;

	// ---- IC generator (parser.hpp:769-972)
	st->seed = 0; P.get("seed", st->seed);
	if (P.get("IC generator", s) && !(s[0] == 'B' || s[0] == 'b')) return fail("only \"IC generator = basic\" is built");
	std::vector<std::string> files = P.items("template file");
	if (files.empty()) return fail("no template file specified");
	for (int i = 0; i < 2; i++) copy_str(st->template_file[i], GEVB_PATH_MAX, files[(size_t) i < files.size() ? (size_t) i : files.size() - 1]);
	if (P.find("mPk file") != NULL) return fail("mPk file: initial conditions from a power spectrum are not built (use a Tk file)");
	if (!P.get("Tk file", s)) return fail("no transfer function file specified (CLASS is not available)");
	copy_str(st->tk_file, GEVB_PATH_MAX, s);
	st->correct_displacement = P.get("correct displacement", s) && (s[0] == 'Y' || s[0] == 'y');
	st->ksphere = P.get("k-domain", s) && (s[0] == 'S' || s[0] == 's');
	{
		std::vector<std::string> t = P.items("tiling factor");
		int last = 1, n = 0;
		for (int i = 0; i < 2; i++)
		{
			if ((size_t) i < t.size() && std::sscanf(t[i].c_str(), "%d", &last) == 1) n++;
			st->tiling[i] = last;
		}
		if (n == 0) st->tiling[0] = st->tiling[1] = 1;
		if (st->tiling[0] <= 0) st->tiling[0] = 1;
		if (st->tiling[1] < 0) st->tiling[1] = 1;
	}
	st->baryon_flag = 2;                                                                       // default: blend (:966-970)
	if (P.get("baryon treatment", s))
	{
		switch (s[0])
		{
			case 'i': case 'I': st->baryon_flag = 0; break;
			case 's': case 'S': st->baryon_flag = 1; break;
			case 'b': case 'B': st->baryon_flag = 2; break;
			case 'h': case 'H': st->baryon_flag = 3; break;
			default: return fail("baryon treatment not supported");
		}
	}
	if (st->baryon_flag == 1 && st->tiling[1] <= 0) st->tiling[1] = 1;
	if (P.get("radiation treatment", s) && (s[0] == 'c' || s[0] == 'C')) return fail("radiation treatment = CLASS is not available");
	if (P.get("fluid treatment", s) && (s[0] == 'c' || s[0] == 'C')) return fail("fluid treatment = CLASS is not available");
	st->z_relax = -2.; P.get("relaxation redshift", st->z_relax);
	st->A_s = 2.215e-9; st->n_s = 0.9619; st->k_pivot = 0.05;                                  // P_SPECTRAL_AMP, P_SPECTRAL_INDEX, P_PIVOT_SCALE (metadata.hpp:74-82)
	P.get("A_s", st->A_s); P.get("n_s", st->n_s); P.get("k_pivot", st->k_pivot);

	// ---- simulation settings (parser.hpp:1138-1262)
	st->vector_flag = 0;                                                                       // VECTOR_PARABOLIC
	if (P.get("vector method", s) && (s[0] == 'e' || s[0] == 'E')) st->vector_flag = 1;
	copy_str(st->basename_generic, sizeof(st->basename_generic), P.get("generic file base", s) ? s : std::string());
	copy_str(st->basename_snapshot, sizeof(st->basename_snapshot), P.get("snapshot file base", s) ? s : std::string("snapshot"));
	copy_str(st->basename_pk, sizeof(st->basename_pk), P.get("Pk file base", s) ? s : std::string("pk"));
	copy_str(st->output_path, GEVB_PATH_MAX, P.get("output path", s) ? s : std::string());
	if (!P.get("boxsize", st->boxsize) || st->boxsize <= 0. || !std::isfinite(st->boxsize)) return fail("simulation box size not set properly");
	if (!P.get("Ngrid", st->ngrid) || st->ngrid < 2) return fail("number of grid points not set properly");
	if (!P.get("Courant factor", st->Cf)) return fail("Courant factor not set");
	if (!P.get("time step limit", st->steplimit)) return fail("time step limit not set");
	if (!P.get("move limit", st->movelimit)) st->movelimit = (double) st->ngrid;
	if (!P.get("initial redshift", st->z_in)) return fail("initial redshift not specified");
	if (st->z_relax < -1.) st->z_relax = st->z_in;
	auto redshifts = [&](const char * name, double * z, int & n)
	{
		n = 0;
		for (const std::string & it : P.items(name)) if (n < GEVB_MAX_OUTPUTS && std::sscanf(it.c_str(), "%lf", &z[n]) == 1) n++;
		for (int i = 1; i < n; i++) for (int j = i; j > 0 && z[j] > z[j - 1]; j--) { const double t = z[j]; z[j] = z[j - 1]; z[j - 1] = t; }   // descending (:1258-1263)
	};
	redshifts("snapshot redshifts", st->z_snapshot, st->num_snapshot);
	redshifts("Pk redshifts", st->z_pk, st->num_pk);
	st->snapshot_mask = field_mask(P, "snapshot outputs");
	st->pk_mask = field_mask(P, "Pk outputs");
	{
		std::vector<std::string> t = P.items("tracer factor");
		for (int i = 0; i < 2; i++)
		{
			st->tracer_factor[i] = 1;
			if ((size_t) i < t.size()) std::sscanf(t[i].c_str(), "%d", &st->tracer_factor[i]);
			if (st->tracer_factor[i] < 1) st->tracer_factor[i] = 1;
		}
	}
	if (!P.get("Pk bins", st->numbins)) st->numbins = 64;
	st->gr_flag = 1;
	if (P.get("gravity theory", s) && (s[0] == 'N' || s[0] == 'n')) st->gr_flag = 0;

	// ---- cosmological parameters (parser.hpp:1607-1745)
	double h = 0.67556;                                                                        // P_HUBBLE (metadata.hpp)
	P.get("h", h);
	if (P.find("m_ncdm") != NULL || (P.get("N_ncdm", tmp) && tmp > 0.)) return fail("ncdm particle species from the IC generator are not built");
	double Omega_g = 0., Omega_ur, Omega_fld = 0., w0 = -1., wa = 0., Omega_b = 0., Omega_cdm = 1.;
	if (P.get("T_cmb", Omega_g))
	{
		Omega_g = Omega_g * Omega_g / h;
		Omega_g = Omega_g * Omega_g * GEVB_C_PLANCK_LAW;                                       // Planck's law
	}
	else if (P.get("omega_g", Omega_g)) Omega_g /= h * h;
	else if (!P.get("Omega_g", Omega_g)) Omega_g = 0.;
	if (P.get("N_ur", Omega_ur)) Omega_ur *= (7. / 8.) * std::pow(4. / 11., 4. / 3.) * Omega_g;
	else if (P.get("N_eff", Omega_ur)) Omega_ur *= (7. / 8.) * std::pow(4. / 11., 4. / 3.) * Omega_g;
	else if (P.get("omega_ur", Omega_ur)) Omega_ur /= h * h;
	else if (!P.get("Omega_ur", Omega_ur)) Omega_ur = 3.046 * (7. / 8.) * std::pow(4. / 11., 4. / 3.) * Omega_g;   // P_N_UR
	const double Omega_rad = Omega_g + Omega_ur;
	if (P.get("omega_fld", Omega_fld)) Omega_fld /= h * h;
	else if (!P.get("Omega_fld", Omega_fld)) Omega_fld = 0.;
	P.get("w0_fld", w0); P.get("wa_fld", wa);
	if (Omega_fld > 0 && w0 == -1.) Omega_fld = 0.;
	if (P.get("omega_b", Omega_b)) Omega_b /= h * h;
	else if (!P.get("Omega_b", Omega_b)) Omega_b = 0.;
	if (P.get("omega_cdm", Omega_cdm)) Omega_cdm /= h * h;
	else if (!P.get("Omega_cdm", Omega_cdm)) Omega_cdm = 1.;
	const double Omega_m = Omega_cdm + Omega_b;
	if (Omega_m <= 0. || Omega_m > 1.) return fail("total matter density out of range");
	if (Omega_rad < 0. || Omega_rad > 1. - Omega_m) return fail("total radiation energy density out of range");
	const double c[11] = {Omega_cdm, Omega_b, Omega_m, 1. - Omega_m - Omega_rad - Omega_fld, Omega_fld, w0, wa, Omega_g, Omega_ur, Omega_rad, h};
	for (int i = 0; i < 11; i++) st->cosmo[i] = c[i];
	return 0;
}
