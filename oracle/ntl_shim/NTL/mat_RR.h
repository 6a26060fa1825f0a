// oracle/ntl_shim/NTL/mat_RR.h - TEST INFRASTRUCTURE ONLY (see RR.h): NTL::mat_RR with SetDims, row access,
// transpose, inv (Gauss-Jordan with partial pivoting, determinant returned like NTL's inv(d, X, A)) and the
// vector * matrix product, which is all common/Remez.cpp:176-213 and common/MinicompRemez.cpp:20-120 need.
#pragma once
#include "vec_RR.h"

namespace NTL
{
	class mat_RR {
	public:
		std::vector<vec_RR> r;
		long cols = 0;
		void SetDims(long n, long m) { r.resize((size_t)n); for (auto &row : r) row.SetLength(m); cols = m; }
		long NumRows() const { return (long)r.size(); }
		long NumCols() const { return cols; }
		vec_RR &operator[](long i) { return r[(size_t)i]; }
		const vec_RR &operator[](long i) const { return r[(size_t)i]; }
		void kill() { r.clear(); cols = 0; }
	};
	inline void transpose(mat_RR &x, const mat_RR &a) {
		mat_RR t; long n = a.NumRows(), m = a.NumCols(); t.SetDims(m, n);
		for (long i = 0; i < n; i++) for (long j = 0; j < m; j++) t[j][i] = a[i][j];
		x = t;
	}
	inline mat_RR transpose(const mat_RR &a) { mat_RR t; transpose(t, a); return t; }
	inline void inv(RR &d, mat_RR &x, const mat_RR &a) {
		long n = a.NumRows();
		if (n != a.NumCols()) { std::cerr << "NTL shim: inv of a non-square matrix\n"; std::abort(); }
		mat_RR w = a, y; y.SetDims(n, n);
		for (long i = 0; i < n; i++) y[i][i] = RR(1);
		RR det(1);
		for (long c = 0; c < n; c++) {
			long p = c;
			for (long i = c + 1; i < n; i++) if (abs(w[i][c]) > abs(w[p][c])) p = i;
			if (IsZero(w[p][c])) { d = RR(); return; }
			if (p != c) { std::swap(w.r[(size_t)p], w.r[(size_t)c]); std::swap(y.r[(size_t)p], y.r[(size_t)c]); det = -det; }
			det = det * w[c][c];
			RR piv = 1 / w[c][c];
			for (long j = 0; j < n; j++) { w[c][j] = w[c][j] * piv; y[c][j] = y[c][j] * piv; }
			for (long i = 0; i < n; i++) {
				if (i == c || IsZero(w[i][c])) continue;
				RR f = w[i][c];
				for (long j = 0; j < n; j++) {
					if (j >= c) w[i][j] = w[i][j] - f * w[c][j];
					y[i][j] = y[i][j] - f * y[c][j];
				}
			}
		}
		d = det; x = y;
	}
	inline void inv(mat_RR &x, const mat_RR &a) { RR d; inv(d, x, a); if (IsZero(d)) { std::cerr << "NTL shim: singular matrix\n"; std::abort(); } }
	inline vec_RR operator*(const vec_RR &v, const mat_RR &a) {   // row vector times matrix
		vec_RR out; long n = a.NumRows(), m = a.NumCols(); out.SetLength(m);
		for (long j = 0; j < m; j++) { RR s; for (long i = 0; i < n; i++) s = s + v[i] * a[i][j]; out[j] = s; }
		return out;
	}
	inline vec_RR operator*(const mat_RR &a, const vec_RR &v) {
		vec_RR out; long n = a.NumRows(), m = a.NumCols(); out.SetLength(n);
		for (long i = 0; i < n; i++) { RR s; for (long j = 0; j < m; j++) s = s + a[i][j] * v[j]; out[i] = s; }
		return out;
	}
	inline void mul(vec_RR &x, const vec_RR &v, const mat_RR &a) { x = v * a; }
}
