// oracle/ntl_shim/NTL/vec_RR.h - TEST INFRASTRUCTURE ONLY (see RR.h): NTL::vec_RR as the reference uses it
// (common/Remez.cpp:180-210: SetLength, operator[], vector * matrix).
#pragma once
#include "RR.h"

namespace NTL
{
	class vec_RR {
	public:
		std::vector<RR> d;
		void SetLength(long n) { d.resize((size_t)n); }
		long length() const { return (long)d.size(); }
		RR &operator[](long i) { return d[(size_t)i]; }
		const RR &operator[](long i) const { return d[(size_t)i]; }
		RR &operator()(long i) { return d[(size_t)(i - 1)]; }
		void kill() { d.clear(); }
	};
	inline void clear(vec_RR &v) { for (auto &x : v.d) x = RR(); }
	inline RR InnerProduct(const vec_RR &a, const vec_RR &b) {
		RR s; long n = std::min(a.length(), b.length()); for (long i = 0; i < n; i++) s = s + a[i] * b[i]; return s;
	}
	inline void InnerProduct(RR &z, const vec_RR &a, const vec_RR &b) { z = InnerProduct(a, b); }
}
