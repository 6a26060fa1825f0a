// oracle/ntl_shim/NTL/RR.h - TEST INFRASTRUCTURE ONLY.
//
// A stand-in for the part of NTL (Shoup's Number Theory Library; the reference links `ntl gmp`,
// /root/reference/cnn_ckks/CMakeLists.txt:57, version not pinned) that the reference's application layers use:
// the arbitrary-precision floating-point class NTL::RR with its free functions.  NTL and the GMP headers are not
// installed in this image and cannot be fetched, so this header lets the reference's own, unmodified sources
// (common/*.cpp, ckks_bootstrapping/*.cpp, comp/*.cpp, cnn/*.cpp) compile - against the reference's SEAL as the
// L2-L4 oracle (oracle/_ref/libcnn_ref.so) and against this repo's seal:: facade as the drop-in proof
// (oracle/_ref/libcnn_dropin.so).  It is a real multi-precision type (sign, binary exponent, up to 1280 mantissa
// bits), because the reference's multi-interval Remez (common/Remez.cpp:176-213) inverts a Chebyshev-Vandermonde
// matrix on 49 intervals of width 2^-10, which needs its 1000 bits (RemezParam.h:14).
//
// Semantics follow NTL's documentation of RR: value = sign * mantissa * 2^exponent, every operation is rounded to
// the current precision (here: truncated to ceil(p/64)+1 limbs, i.e. at least p+1 bits - a little more accurate
// than NTL, never less), precision is per thread with NTL's default of 150 bits.
// $NTL_SHIM_MAX_PREC caps the precision (tests use it to keep the Remez quick).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <iostream>
#include <mutex>
#include <unordered_map>
#include <string>
#include <type_traits>
#include <vector>

namespace NTL
{
	class ZZ;

	class RR {
	public:
		static constexpr int MAXL = 20;
		int sgn = 0;          // -1, 0, +1
		long e = 0;           // value = sgn * 0.m[0]m[1]... * 2^e, top bit of m[0] set when sgn != 0
		uint64_t m[MAXL];     // big-endian limbs

		RR() { std::memset(m, 0, sizeof m); }
		template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
		RR(T v) { std::memset(m, 0, sizeof m); assign(v); }
		template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
		RR &operator=(T v) { std::memset(m, 0, sizeof m); sgn = 0; e = 0; assign(v); return *this; }

		// ---- precision (thread local, as in NTL) ----
		static long &prec_ref() { static thread_local long p = 150; return p; }
		static long &oprec_ref() { static thread_local long p = 10; return p; }
		static void SetPrecision(long p) {
			static const long cap = [] { const char *s = std::getenv("NTL_SHIM_MAX_PREC"); return s ? std::atol(s) : 0L; }();
			if (cap > 0 && p > cap) p = cap;
			if (p < 53) p = 53;
			if (p > 64L * (MAXL - 1)) p = 64L * (MAXL - 1);
			prec_ref() = p;
		}
		static long precision() { return prec_ref(); }
		static void SetOutputPrecision(long p) { oprec_ref() = p < 1 ? 1 : p; }
		static long OutputPrecision() { return oprec_ref(); }
		static int L() { int l = (int)((prec_ref() + 63) / 64) + 1; return l > MAXL ? MAXL : l; }

		bool is_zero() const { return sgn == 0; }

	private:
		void assign(double v) {
			if (v == 0 || v != v) return;
			int ex; double f = std::frexp(std::fabs(v), &ex);   // f in [0.5, 1)
			m[0] = (uint64_t)std::ldexp(f, 53) << 11; e = ex; sgn = v < 0 ? -1 : 1;
		}
		void assign(float v) { assign((double)v); }
		void assign(long double v) { assign((double)v); }
		template <class T> typename std::enable_if<std::is_integral<T>::value>::type assign(T v) {
			if (v == 0) return;
			uint64_t a; bool neg = false;
			if (std::is_signed<T>::value && v < 0) { neg = true; a = (uint64_t)(-(int64_t)v); } else a = (uint64_t)v;
			int z = __builtin_clzll(a);
			m[0] = a << z; e = 64 - z; sgn = neg ? -1 : 1;
		}
	};

	namespace rr_detail
	{
		typedef unsigned __int128 u128;
		constexpr int W = RR::MAXL + 2;

		inline int cmp_mag(const RR &a, const RR &b) {   // both non-zero
			if (a.e != b.e) return a.e < b.e ? -1 : 1;
			for (int i = 0; i < RR::MAXL; i++) if (a.m[i] != b.m[i]) return a.m[i] < b.m[i] ? -1 : 1;
			return 0;
		}
		// t[0..n) = a >> s (s bits, big-endian limbs), a has na limbs
		inline void shr_into(uint64_t *t, int n, const uint64_t *a, int na, long s) {
			long ls = s / 64; int bs = (int)(s % 64);
			for (int i = 0; i < n; i++) {
				long src = i - ls;
				uint64_t hi = (src >= 0 && src < na) ? a[src] : 0, hh = (src - 1 >= 0 && src - 1 < na) ? a[src - 1] : 0;
				t[i] = bs ? (hi >> bs) | (hh << (64 - bs)) : hi;
			}
		}
		// normalise t[0..n) (shift left so that the top bit is set), store into r with exponent ex; returns false if zero
		inline void norm_store(RR &r, uint64_t *t, int n, long ex, int sgn) {
			int lz = 0; while (lz < n && t[lz] == 0) lz++;
			std::memset(r.m, 0, sizeof r.m);
			if (lz == n) { r.sgn = 0; r.e = 0; return; }
			int bz = __builtin_clzll(t[lz]);
			int L = RR::L();
			for (int i = 0; i < L; i++) {
				int s = i + lz;
				uint64_t hi = s < n ? t[s] : 0, lo = s + 1 < n ? t[s + 1] : 0;
				r.m[i] = bz ? (hi << bz) | (lo >> (64 - bz)) : hi;
			}
			r.e = ex - 64L * lz - bz; r.sgn = sgn;
		}
		inline void add_signed(RR &r, const RR &a, const RR &b, int bsgn) {   // r = a + bsgn*b
			if (bsgn == 0 || b.sgn == 0) { RR t = a; r = t; return; }
			int sb = b.sgn * bsgn;
			if (a.sgn == 0) { RR t = b; t.sgn = sb; r = t; return; }
			const RR *x = &a, *y = &b; int sx = a.sgn, sy = sb;
			if (cmp_mag(a, b) < 0) { x = &b; y = &a; sx = sb; sy = a.sgn; }
			int L = RR::L(), n = L + 2;          // one limb of headroom in front, one guard limb behind
			long d = x->e - y->e;
			uint64_t tx[W + 1], ty[W + 1];
			shr_into(tx, n, x->m, RR::MAXL, 64);
			if (d >= 64L * (L + 1)) { norm_store(r, tx, n, x->e + 64, sx); return; }
			shr_into(ty, n, y->m, RR::MAXL, 64 + d);
			if (sx == sy) {
				unsigned c = 0;
				for (int i = n - 1; i >= 0; i--) { u128 s = (u128)tx[i] + ty[i] + c; tx[i] = (uint64_t)s; c = (unsigned)(s >> 64); }
			} else {
				unsigned bw = 0;
				for (int i = n - 1; i >= 0; i--) { u128 s = (u128)tx[i] - ty[i] - bw; tx[i] = (uint64_t)s; bw = (unsigned)((s >> 64) & 1); }
			}
			norm_store(r, tx, n, x->e + 64, sx);
		}
		inline void mul(RR &r, const RR &a, const RR &b) {
			if (a.sgn == 0 || b.sgn == 0) { r = RR(); return; }
			int L = RR::L();
			uint64_t t[W + 1];
			// columns k = i + j; products reach limbs k (high) and k+1 (low); keep limbs 0..L+1
			u128 acc = 0; uint64_t over = 0;
			for (int k = L; k >= 0; k--) {
				int i0 = k - (L - 1) > 0 ? k - (L - 1) : 0, i1 = k < L - 1 ? k : L - 1;
				for (int i = i0; i <= i1; i++) {
					u128 p = (u128)a.m[i] * b.m[k - i];
					acc += p; if (acc < p) over++;
				}
				t[k + 1] = (uint64_t)acc; acc = (acc >> 64) | ((u128)over << 64); over = 0;
			}
			t[0] = (uint64_t)acc;
			norm_store(r, t, L + 2, a.e + b.e, a.sgn * b.sgn);
		}
		inline void div_small(RR &r, const RR &a, uint64_t d, int dsgn) {
			if (a.sgn == 0) { r = RR(); return; }
			int L = RR::L(), n = L + 2;
			uint64_t t[W + 1]; u128 rem = 0;
			for (int i = 0; i < n; i++) {
				uint64_t limb = i < RR::MAXL ? a.m[i] : 0;
				u128 cur = (rem << 64) | limb; t[i] = (uint64_t)(cur / d); rem = cur % d;
			}
			norm_store(r, t, n, a.e, a.sgn * dsgn);
		}
		inline double mant_to_double(const RR &a) { return std::ldexp((double)(a.m[0] >> 11), -53); }  // [0.5,1), truncated
		inline void inv(RR &r, const RR &b) {   // Newton: x <- x + x(1 - b x)
			if (b.sgn == 0) { std::cerr << "NTL shim: division by zero\n"; std::abort(); }
			RR f = b; f.e = 0; f.sgn = 1;                 // f in [0.5,1)
			RR x(1.0 / mant_to_double(f));
			RR one(1), t, u;
			int iters = 0; long bits = 50, need = 64L * RR::L() + 8;
			while (bits < need) { bits *= 2; iters++; }
			for (int i = 0; i < iters + 1; i++) { mul(t, f, x); add_signed(u, one, t, -1); mul(t, x, u); add_signed(x, x, t, +1); }
			x.e -= b.e; x.sgn = b.sgn; r = x;
		}
	}

	// ---- arithmetic ----
	inline RR operator-(const RR &a) { RR r = a; r.sgn = -r.sgn; return r; }
	inline RR operator+(const RR &a) { return a; }
	inline RR operator+(const RR &a, const RR &b) { RR r; rr_detail::add_signed(r, a, b, +1); return r; }
	inline RR operator-(const RR &a, const RR &b) { RR r; rr_detail::add_signed(r, a, b, -1); return r; }
	inline RR operator*(const RR &a, const RR &b) { RR r; rr_detail::mul(r, a, b); return r; }
	inline RR operator/(const RR &a, const RR &b) { RR i, r; rr_detail::inv(i, b); rr_detail::mul(r, a, i); return r; }
	inline int compare(const RR &a, const RR &b) {
		if (a.sgn != b.sgn) return a.sgn < b.sgn ? -1 : 1;
		if (a.sgn == 0) return 0;
		return a.sgn * rr_detail::cmp_mag(a, b);
	}
	inline bool operator==(const RR &a, const RR &b) { return compare(a, b) == 0; }
	inline bool operator!=(const RR &a, const RR &b) { return compare(a, b) != 0; }
	inline bool operator<(const RR &a, const RR &b) { return compare(a, b) < 0; }
	inline bool operator<=(const RR &a, const RR &b) { return compare(a, b) <= 0; }
	inline bool operator>(const RR &a, const RR &b) { return compare(a, b) > 0; }
	inline bool operator>=(const RR &a, const RR &b) { return compare(a, b) >= 0; }

#define NTL_SHIM_ARITH template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
	NTL_SHIM_ARITH inline RR operator+(const RR &a, T b) { return a + RR(b); }
	NTL_SHIM_ARITH inline RR operator+(T a, const RR &b) { return RR(a) + b; }
	NTL_SHIM_ARITH inline RR operator-(const RR &a, T b) { return a - RR(b); }
	NTL_SHIM_ARITH inline RR operator-(T a, const RR &b) { return RR(a) - b; }
	NTL_SHIM_ARITH inline RR operator*(const RR &a, T b) { return a * RR(b); }
	NTL_SHIM_ARITH inline RR operator*(T a, const RR &b) { return RR(a) * b; }
	NTL_SHIM_ARITH inline RR operator/(T a, const RR &b) { return RR(a) / b; }
	NTL_SHIM_ARITH inline RR operator/(const RR &a, T b) {
		if (std::is_integral<T>::value) {
			if (b == 0) { std::cerr << "NTL shim: division by zero\n"; std::abort(); }
			bool neg = b < 0; uint64_t d = neg ? (uint64_t)(-(int64_t)b) : (uint64_t)b;
			RR r; rr_detail::div_small(r, a, d, neg ? -1 : 1); return r;
		}
		return a / RR(b);
	}
	NTL_SHIM_ARITH inline bool operator==(const RR &a, T b) { return compare(a, RR(b)) == 0; }
	NTL_SHIM_ARITH inline bool operator!=(const RR &a, T b) { return compare(a, RR(b)) != 0; }
	NTL_SHIM_ARITH inline bool operator<(const RR &a, T b) { return compare(a, RR(b)) < 0; }
	NTL_SHIM_ARITH inline bool operator<=(const RR &a, T b) { return compare(a, RR(b)) <= 0; }
	NTL_SHIM_ARITH inline bool operator>(const RR &a, T b) { return compare(a, RR(b)) > 0; }
	NTL_SHIM_ARITH inline bool operator>=(const RR &a, T b) { return compare(a, RR(b)) >= 0; }
	NTL_SHIM_ARITH inline bool operator==(T a, const RR &b) { return compare(RR(a), b) == 0; }
	NTL_SHIM_ARITH inline bool operator!=(T a, const RR &b) { return compare(RR(a), b) != 0; }
	NTL_SHIM_ARITH inline bool operator<(T a, const RR &b) { return compare(RR(a), b) < 0; }
	NTL_SHIM_ARITH inline bool operator<=(T a, const RR &b) { return compare(RR(a), b) <= 0; }
	NTL_SHIM_ARITH inline bool operator>(T a, const RR &b) { return compare(RR(a), b) > 0; }
	NTL_SHIM_ARITH inline bool operator>=(T a, const RR &b) { return compare(RR(a), b) >= 0; }

	inline RR &operator+=(RR &a, const RR &b) { a = a + b; return a; }
	inline RR &operator-=(RR &a, const RR &b) { a = a - b; return a; }
	inline RR &operator*=(RR &a, const RR &b) { a = a * b; return a; }
	inline RR &operator/=(RR &a, const RR &b) { a = a / b; return a; }
	NTL_SHIM_ARITH inline RR &operator+=(RR &a, T b) { a = a + b; return a; }
	NTL_SHIM_ARITH inline RR &operator-=(RR &a, T b) { a = a - b; return a; }
	NTL_SHIM_ARITH inline RR &operator*=(RR &a, T b) { a = a * b; return a; }
	NTL_SHIM_ARITH inline RR &operator/=(RR &a, T b) { a = a / b; return a; }
	inline RR &operator++(RR &a) { a = a + 1; return a; }
	inline RR &operator--(RR &a) { a = a - 1; return a; }

	// procedural forms
	inline void add(RR &z, const RR &a, const RR &b) { z = a + b; }
	inline void sub(RR &z, const RR &a, const RR &b) { z = a - b; }
	inline void mul(RR &z, const RR &a, const RR &b) { z = a * b; }
	inline void div(RR &z, const RR &a, const RR &b) { z = a / b; }
	inline void negate(RR &z, const RR &a) { z = -a; }
	inline void clear(RR &z) { z = RR(); }
	inline void set(RR &z) { z = RR(1); }
	inline bool IsZero(const RR &a) { return a.sgn == 0; }
	inline bool IsOne(const RR &a) { return a == 1; }
	inline long sign(const RR &a) { return a.sgn; }
	inline RR abs(const RR &a) { RR r = a; if (r.sgn < 0) r.sgn = 1; return r; }
	inline RR fabs(const RR &a) { return abs(a); }

	// ---- conversions ----
	inline double to_double(const RR &a) {
		if (a.sgn == 0) return 0.0;
		// round to nearest on 53 bits
		uint64_t top = a.m[0] >> 11, rest = a.m[0] & 0x7FF;
		bool up = rest > 0x400 || (rest == 0x400 && ((top & 1) || a.m[1] != 0)) ;
		if (rest == 0x400 && !up) { for (int i = 1; i < RR::MAXL && !up; i++) up = a.m[i] != 0; }
		double f = std::ldexp((double)top + (up ? 1.0 : 0.0), -53);
		return a.sgn * std::ldexp(f, (int)std::max<long>(std::min<long>(a.e, 100000), -100000));
	}
	inline float to_float(const RR &a) { return (float)to_double(a); }
	NTL_SHIM_ARITH inline RR to_RR(T v) { return RR(v); }
	inline RR to_RR(const RR &v) { return v; }
	inline RR to_RR(const char *s);
	inline void conv(RR &z, double a) { z = RR(a); }
	inline void conv(RR &z, long a) { z = RR(a); }
	inline void conv(RR &z, int a) { z = RR(a); }
	inline void conv(RR &z, const RR &a) { z = a; }
	inline void conv(double &z, const RR &a) { z = to_double(a); }
	template <class T> inline T conv(const RR &a);
	template <> inline double conv<double>(const RR &a) { return to_double(a); }
	inline RR power2_RR(long k) { RR r(1); r.e += k; return r; }
	inline void power2(RR &z, long k) { z = power2_RR(k); }
	inline RR MakeRR(long mant, long ex) { RR r(mant); if (r.sgn) r.e += ex; return r; }

	// ---- rounding to integers ----
	inline RR trunc(const RR &a) {
		if (a.sgn == 0 || a.e <= 0) return RR();
		RR r = a; long keep = a.e;
		for (int i = 0; i < RR::MAXL; i++) {
			long lo = 64L * i;
			if (keep <= lo) r.m[i] = 0; else if (keep < lo + 64) r.m[i] &= ~uint64_t(0) << (64 - (keep - lo));
		}
		return r;
	}
	inline RR floor(const RR &a) { RR t = trunc(a); if (a.sgn < 0 && t != a) t = t - 1; return t; }
	inline RR ceil(const RR &a) { RR t = trunc(a); if (a.sgn > 0 && t != a) t = t + 1; return t; }
	inline RR round(const RR &a) {   // nearest integer, ties to even (NTL's rule)
		RR f = floor(a), d = a - f, half(0.5);
		int c = compare(d, half);
		if (c < 0) return f;
		if (c > 0) return f + 1;
		RR h = f / 2; return floor(h) == h ? f : f + 1;
	}
	inline long to_long(const RR &a) {   // floor, as NTL's conv(long&, RR)
		RR f = floor(a);
		if (f.sgn == 0) return 0;
		if (f.e > 63) return f.sgn > 0 ? INT64_MAX : INT64_MIN;
		uint64_t v = f.m[0] >> (64 - f.e);
		return f.sgn > 0 ? (long)v : -(long)v;
	}
	inline int to_int(const RR &a) { return (int)to_long(a); }
	inline void conv(long &z, const RR &a) { z = to_long(a); }

	// ---- powers, roots ----
	inline RR power(const RR &a, long n) {
		if (n < 0) return 1 / power(a, -n);
		RR r(1), b = a;
		while (n) { if (n & 1) r = r * b; n >>= 1; if (n) b = b * b; }
		return r;
	}
	inline void power(RR &z, const RR &a, long n) { z = power(a, n); }
	inline RR sqr(const RR &a) { return a * a; }
	inline RR inv(const RR &a) { return 1 / a; }
	inline RR sqrt(const RR &a) {
		if (a.sgn == 0) return RR();
		if (a.sgn < 0) { std::cerr << "NTL shim: sqrt of a negative number\n"; std::abort(); }
		// f = mantissa * 2^(0 or 1) so that the remaining exponent is even; y -> 1/sqrt(f) by Newton, sqrt = f*y
		RR f = a; long ex = a.e; f.e = 0; if (ex & 1) { f.e = 1; ex -= 1; }
		double fd = rr_detail::mant_to_double(f) * (f.e ? 2.0 : 1.0);
		RR y(1.0 / std::sqrt(fd)), three(3);
		long bits = 50, need = 64L * RR::L() + 8; int iters = 1;
		while (bits < need) { bits *= 2; iters++; }
		for (int i = 0; i < iters; i++) y = y * (three - f * y * y) / 2;
		RR r = f * y; r.e += ex / 2; return r;
	}
	inline RR SqrRoot(const RR &a) { return sqrt(a); }

	// ---- pi, sin, cos, exp, log ----
	namespace rr_detail
	{
		inline RR atan_inv(long k) {   // atan(1/k) = sum (-1)^i / ((2i+1) k^(2i+1))
			RR term = RR(1) / k, sum = term; long k2 = k * k, need = 64L * RR::L() + 8;
			for (long i = 1;; i++) {
				term = term / k2; RR t = term / (2 * i + 1);
				if (t.sgn == 0 || t.e < sum.e - need) break;
				sum = (i & 1) ? sum - t : sum + t;
			}
			return sum;
		}
	}
	inline RR ComputePi_RR() {
		static thread_local long have = 0; static thread_local RR pi;
		if (have != RR::precision()) { pi = 16 * rr_detail::atan_inv(5) - 4 * rr_detail::atan_inv(239); have = RR::precision(); }
		return pi;
	}
	inline void ComputePi(RR &z) { z = ComputePi_RR(); }
	namespace rr_detail
	{
		// sin and cos of |y| <= pi/4 + eps: Taylor series of y / 2^h, then h double-angle steps
		// (s, c) <- (2 s c, (c - s)(c + s)); the working precision is raised by 2 limbs for the error the
		// doublings amplify
		inline void sincos_small(const RR &y, RR &s, RR &c) {
			if (y.sgn == 0) { s = RR(); c = RR(1); return; }
			const long old = RR::precision();
			const int h = 24;
			RR::SetPrecision(old + 128);
			long need = 64L * RR::L() + 8;
			RR z = y; z.e -= h;
			RR z2 = z * z, term(1), ts = z, cc(1), ss = z;
			for (long i = 1;; i++) {
				term = term * z2 / ((2 * i - 1) * (2 * i));      // z^(2i)/(2i)!
				ts = ts * z2 / ((2 * i) * (2 * i + 1));         // z^(2i+1)/(2i+1)!
				bool done = (term.sgn == 0 || term.e < -need) && (ts.sgn == 0 || ts.e < z.e - need);
				if (i & 1) { cc = cc - term; ss = ss - ts; } else { cc = cc + term; ss = ss + ts; }
				if (done) break;
			}
			for (int i = 0; i < h; i++) {
				RR s2 = ss * cc; s2.e += (s2.sgn ? 1 : 0);
				RR c2 = (cc - ss) * (cc + ss);
				ss = s2; cc = c2;
			}
			RR::SetPrecision(old);
			RR one(1); s = ss * one; c = cc * one;
		}
		// memo of (sin, cos) by argument: the reference's Remez scans the same grid of points in every iteration
		// (common/Remez.cpp:240-246), from threads it recreates per iteration, hence a process-wide table
		struct SinCosMemo {
			struct Entry { RR x, s, c; long prec; };
			static constexpr int SHARDS = 64;
			std::mutex mu[SHARDS];
			std::unordered_multimap<uint64_t, Entry> map[SHARDS];
			std::atomic<size_t> count{0};
			static SinCosMemo &get() { static SinCosMemo m; return m; }
			static uint64_t hash(const RR &x) {
				uint64_t h = 1469598103934665603ull ^ (uint64_t)x.e ^ ((uint64_t)(x.sgn + 1) << 62);
				for (int i = 0; i < RR::MAXL; i++) { h ^= x.m[i]; h *= 1099511628211ull; }
				return h;
			}
			bool find(const RR &x, RR &s, RR &c) {
				uint64_t h = hash(x); int sh = (int)(h % SHARDS);
				std::lock_guard<std::mutex> g(mu[sh]);
				auto r = map[sh].equal_range(h);
				for (auto it = r.first; it != r.second; ++it)
					if (it->second.prec == RR::precision() && compare(it->second.x, x) == 0) { s = it->second.s; c = it->second.c; return true; }
				return false;
			}
			void put(const RR &x, const RR &s, const RR &c) {
				if (count.load() > 400000) return;      // ~350 MB cap
				uint64_t h = hash(x); int sh = (int)(h % SHARDS);
				std::lock_guard<std::mutex> g(mu[sh]);
				map[sh].emplace(h, Entry{ x, s, c, RR::precision() });
				count++;
			}
		};
		inline void sincos(const RR &x, RR &s, RR &c) {
			if (x.sgn == 0) { s = RR(); c = RR(1); return; }
			if (SinCosMemo::get().find(x, s, c)) return;
			RR pi = ComputePi_RR(), half_pi = pi / 2;
			RR q = round(x / half_pi), y = x - q * half_pi;
			RR q4 = q - 4 * floor(q / 4); long k = to_long(q4) & 3;
			RR ss, cc; sincos_small(y, ss, cc);
			switch (k) {
			case 0: s = ss; c = cc; break;
			case 1: s = cc; c = -ss; break;
			case 2: s = -ss; c = -cc; break;
			default: s = -cc; c = ss; break;
			}
			SinCosMemo::get().put(x, s, c);
		}
	}
	inline RR sin(const RR &x) { RR s, c; rr_detail::sincos(x, s, c); return s; }
	inline RR cos(const RR &x) { RR s, c; rr_detail::sincos(x, s, c); return c; }
	inline RR exp(const RR &x) {
		if (x.sgn == 0) return RR(1);
		// x = k ln2 + r is avoided (no ln2 yet): halve until |r| < 2^-8, Taylor, square back
		long halvings = x.e + 8 > 0 ? x.e + 8 : 0;
		RR r = x; r.e -= halvings;
		long need = 64L * RR::L() + 8 + halvings;
		RR sum(1), term(1);
		for (long i = 1;; i++) { term = term * r / i; if (term.sgn == 0 || term.e < -need) break; sum = sum + term; }
		for (long i = 0; i < halvings; i++) sum = sum * sum;
		return sum;
	}
	inline RR log(const RR &x) {
		if (x.sgn <= 0) { std::cerr << "NTL shim: log of a non-positive number\n"; std::abort(); }
		// Newton on exp from a double start: y <- y + (x exp(-y) - 1) ... quadratic
		RR f = x; f.e = 0;
		double y0 = std::log(rr_detail::mant_to_double(f)) + (double)x.e * 0.6931471805599453;
		RR y(y0);
		long bits = 45, need = 64L * RR::L() + 8;
		while (bits < need) { RR ey = exp(-y); y = y + (x * ey - 1); bits *= 2; }
		RR ey = exp(-y); y = y + (x * ey - 1);
		return y;
	}
	inline RR pow(const RR &a, const RR &b) {
		if (floor(b) == b && b.e < 40) return power(a, to_long(b));
		return exp(b * log(a));
	}
	NTL_SHIM_ARITH inline RR pow(const RR &a, T b) { return pow(a, RR(b)); }
	NTL_SHIM_ARITH inline RR pow(T a, const RR &b) { return pow(RR(a), b); }
	inline void pow(RR &z, const RR &a, const RR &b) { z = pow(a, b); }
	inline RR expm1(const RR &x) { return exp(x) - 1; }
	inline RR log1p(const RR &x) { return log(1 + x); }
	inline RR log10(const RR &x) { return log(x) / log(RR(10)); }

	// ---- decimal input / output ----
	namespace rr_detail
	{
		// non-negative integer-valued RR -> decimal digits
		inline std::string int_digits(const RR &a) {
			if (a.sgn == 0 || a.e <= 0) return "0";
			int n = (int)((a.e + 63) / 64);
			std::vector<uint64_t> v(n, 0);             // big-endian integer of a.e bits
			long sh = 64L * n - a.e;                    // a.m as a fraction 0.m -> integer = top a.e bits
			std::vector<uint64_t> src(n + 1, 0);
			for (int i = 0; i < n && i < RR::MAXL; i++) src[i] = a.m[i];
			for (int i = 0; i < n; i++) {               // v = src >> sh  (sh < 64), aligned to n limbs
				uint64_t hi = src[i], hh = i ? src[i - 1] : 0;
				v[i] = sh ? (hi >> sh) | (hh << (64 - sh)) : hi;
			}
			std::string out;
			for (;;) {
				bool zero = true; u128 rem = 0;
				for (int i = 0; i < n; i++) {
					u128 cur = (rem << 64) | v[i]; v[i] = (uint64_t)(cur / 10000000000000000000ULL); rem = cur % 10000000000000000000ULL;
					if (v[i]) zero = false;
				}
				uint64_t chunk = (uint64_t)rem;
				for (int d = 0; d < 19; d++) { out.push_back((char)('0' + chunk % 10)); chunk /= 10; if (zero && chunk == 0) break; }
				if (zero) break;
			}
			while (out.size() > 1 && out.back() == '0') out.pop_back();
			std::reverse(out.begin(), out.end());
			return out;
		}
	}
	inline std::ostream &operator<<(std::ostream &os, const RR &a) {
		if (a.sgn == 0) return os << "0";
		long P = RR::OutputPrecision();
		long old = RR::precision();
		RR::SetPrecision(old + 64);
		RR x = abs(a);
		long d10 = (long)std::floor((double)(x.e - 1) * 0.30102999566398120);   // 10^d10 <= x roughly
		RR scaled = x * power(RR(10), P - 1 - d10);
		RR lim = power(RR(10), P);
		RR r = floor(scaled + RR(0.5));
		if (r >= lim) { d10++; scaled = x * power(RR(10), P - 1 - d10); r = floor(scaled + RR(0.5)); }
		else if (r < lim / 10) { d10--; scaled = x * power(RR(10), P - 1 - d10); r = floor(scaled + RR(0.5)); if (r >= lim) { d10++; r = lim / 10; } }
		std::string dg = rr_detail::int_digits(r);     // P digits
		RR::SetPrecision(old);
		while (dg.size() > 1 && dg.back() == '0') dg.pop_back();
		std::string s = a.sgn < 0 ? "-" : "";
		long point = d10 + 1;                          // digits before the decimal point
		if (point > 0 && point <= 40) {
			if ((long)dg.size() <= point) s += dg + std::string(point - dg.size(), '0');
			else s += dg.substr(0, point) + "." + dg.substr(point);
		} else if (point <= 0 && point > -10) {
			s += "0." + std::string(-point, '0') + dg;
		} else {
			s += "0." + dg + "e" + std::to_string(point);
		}
		return os << s;
	}
	inline bool rr_parse(const std::string &tok, RR &out) {
		size_t i = 0; int sg = 1;
		if (i < tok.size() && (tok[i] == '+' || tok[i] == '-')) { if (tok[i] == '-') sg = -1; i++; }
		long old = RR::precision(); RR::SetPrecision(old + 64);
		RR val; long frac = 0; bool any = false, seen_point = false;
		// accumulate 18 digits at a time
		uint64_t chunk = 0; int cd = 0;
		auto flush = [&]() { if (cd) { val = val * power(RR(10), cd) + RR(chunk); chunk = 0; cd = 0; } };
		for (; i < tok.size(); i++) {
			char ch = tok[i];
			if (ch >= '0' && ch <= '9') { chunk = chunk * 10 + (uint64_t)(ch - '0'); cd++; any = true; if (seen_point) frac++; if (cd == 18) flush(); }
			else if (ch == '.' && !seen_point) seen_point = true;
			else break;
		}
		flush();
		long ex = 0;
		if (i < tok.size() && (tok[i] == 'e' || tok[i] == 'E')) { ex = std::atol(tok.c_str() + i + 1); }
		if (!any) { RR::SetPrecision(old); return false; }
		long p10 = ex - frac;
		if (p10 > 0) val = val * power(RR(10), p10); else if (p10 < 0) val = val / power(RR(10), -p10);
		RR::SetPrecision(old);
		RR one(1); val = val * one;    // round to the caller's precision
		if (sg < 0) val.sgn = -val.sgn;
		out = val; return true;
	}
	inline std::istream &operator>>(std::istream &is, RR &a) {
		std::string tok;
		if (!(is >> tok)) return is;
		if (!rr_parse(tok, a)) is.setstate(std::ios::failbit);
		return is;
	}
	inline RR to_RR(const char *s) { RR r; rr_parse(s, r); return r; }
	inline void conv(RR &z, const char *s) { z = to_RR(s); }
#undef NTL_SHIM_ARITH
}
#include "ZZ.h"   // NTL's RR.h brings ZZ with it (common/MinicompFunc.h:9,19 relies on that)
