// oracle/ntl_shim/NTL/ZZ.h - TEST INFRASTRUCTURE ONLY (see RR.h).  The reference touches NTL::ZZ in two helpers
// (common/func.cpp:8-11 RoundToZZ + parity, common/MinicompFunc.cpp:7-10 centred remainder); a 128-bit integer
// is enough for both.
#pragma once
#include "RR.h"

namespace NTL
{
	class ZZ {
	public:
		__int128 v = 0;
		ZZ() {}
		template <class T, class = typename std::enable_if<std::is_integral<T>::value>::type> ZZ(T x) : v(x) {}
		explicit ZZ(__int128 x, int) : v(x) {}
	};
	inline ZZ mk(__int128 x) { return ZZ(x, 0); }
	inline ZZ operator+(const ZZ &a, const ZZ &b) { return mk(a.v + b.v); }
	inline ZZ operator-(const ZZ &a, const ZZ &b) { return mk(a.v - b.v); }
	inline ZZ operator-(const ZZ &a) { return mk(-a.v); }
	inline ZZ operator*(const ZZ &a, const ZZ &b) { return mk(a.v * b.v); }
	inline ZZ operator/(const ZZ &a, const ZZ &b) {   // floor division, as NTL
		__int128 q = a.v / b.v, r = a.v % b.v; if (r != 0 && ((r < 0) != (b.v < 0))) q -= 1; return mk(q);
	}
	inline ZZ operator%(const ZZ &a, const ZZ &b) { __int128 r = a.v % b.v; if (r != 0 && ((r < 0) != (b.v < 0))) r += b.v; return mk(r); }
	inline long operator%(const ZZ &a, long b) { __int128 r = a.v % b; if (r != 0 && ((r < 0) != (b < 0))) r += b; return (long)r; }
	inline ZZ operator/(const ZZ &a, long b) { return a / ZZ(b); }
	inline bool operator==(const ZZ &a, const ZZ &b) { return a.v == b.v; }
	inline bool operator!=(const ZZ &a, const ZZ &b) { return a.v != b.v; }
	inline bool operator<(const ZZ &a, const ZZ &b) { return a.v < b.v; }
	inline bool operator<=(const ZZ &a, const ZZ &b) { return a.v <= b.v; }
	inline bool operator>(const ZZ &a, const ZZ &b) { return a.v > b.v; }
	inline bool operator>=(const ZZ &a, const ZZ &b) { return a.v >= b.v; }
	inline long to_long(const ZZ &a) { return (long)a.v; }
	inline ZZ to_ZZ(long a) { return ZZ(a); }
	inline ZZ RoundToZZ(const RR &a) {
		RR r = round(a);
		if (r.sgn == 0) return ZZ();
		if (r.e > 126) { std::cerr << "NTL shim: RoundToZZ beyond 126 bits\n"; std::abort(); }
		unsigned __int128 m = ((unsigned __int128)r.m[0] << 64) | r.m[1];
		__int128 v = (__int128)(m >> (128 - r.e));
		return mk(r.sgn < 0 ? -v : v);
	}
	inline ZZ FloorToZZ(const RR &a) { return RoundToZZ(floor(a)); }
	inline ZZ TruncToZZ(const RR &a) { return RoundToZZ(trunc(a)); }
	inline RR to_RR(const ZZ &a) {
		__int128 v = a.v; bool neg = v < 0; unsigned __int128 u = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
		RR hi((uint64_t)(u >> 64)), lo((uint64_t)u);
		RR r = hi * power2_RR(64) + lo; return neg ? -r : r;
	}
	inline std::ostream &operator<<(std::ostream &os, const ZZ &a) { return os << to_RR(a); }
}
