// hostmath.h - host-side (setup-time) number theory for the B200 CKKS engine.
//
// Everything here runs once at context creation: prime search, minimal primitive roots,
// twiddle / Galois tables, RNS constants.  Semantics follow the reference's SEAL 3.6.6
// (cited per function) so that the device tables - and therefore every NTT-form residue -
// are identical to the reference's; the code itself is written from scratch on __int128.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace bk
{
    typedef unsigned __int128 u128;

    inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q)
    {
        return (uint64_t)((u128)a * b % q);
    }
    inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t q)
    {
        uint64_t s = a + b;
        return s >= q ? s - q : s;
    }
    inline uint64_t submod(uint64_t a, uint64_t b, uint64_t q)
    {
        return a >= b ? a - b : a + q - b;
    }
    inline uint64_t powmod(uint64_t a, uint64_t e, uint64_t q)
    {
        uint64_t r = 1 % q;
        a %= q;
        while (e)
        {
            if (e & 1)
                r = mulmod(r, a, q);
            a = mulmod(a, a, q);
            e >>= 1;
        }
        return r;
    }
    inline uint64_t invmod(uint64_t a, uint64_t q) // q prime
    {
        return powmod(a, q - 2, q);
    }
    // Shoup quotient floor(w * 2^64 / q)  (uintarithsmallmod.h:255-285, MultiplyUIntModOperand)
    inline uint64_t shoup(uint64_t w, uint64_t q)
    {
        return (uint64_t)(((u128)w << 64) / q);
    }

    inline uint32_t bitrev(uint32_t x, int bits)
    {
        uint32_t r = 0;
        for (int i = 0; i < bits; i++)
        {
            r = (r << 1) | (x & 1);
            x >>= 1;
        }
        return r;
    }

    // Deterministic Miller-Rabin for 64-bit integers (replaces Modulus::is_prime, modulus.cpp).
    inline bool is_prime_u64(uint64_t n)
    {
        if (n < 2)
            return false;
        static const uint64_t small[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
        for (uint64_t p : small)
        {
            if (n == p)
                return true;
            if (n % p == 0)
                return false;
        }
        uint64_t d = n - 1;
        int r = 0;
        while (!(d & 1))
        {
            d >>= 1;
            r++;
        }
        for (uint64_t a : small)
        {
            uint64_t x = powmod(a, d, n);
            if (x == 1 || x == n - 1)
                continue;
            bool comp = true;
            for (int i = 1; i < r; i++)
            {
                x = mulmod(x, x, n);
                if (x == n - 1)
                {
                    comp = false;
                    break;
                }
            }
            if (comp)
                return false;
        }
        return true;
    }

    // get_primes (util/numth.cpp:277-320): descending scan from 2^bits - 2N + 1 in steps of 2N.
    inline std::vector<uint64_t> get_primes(size_t ntt_size, int bit_size, size_t count)
    {
        if (bit_size < 2 || bit_size > 61)
            throw std::invalid_argument("bit_size is invalid");
        std::vector<uint64_t> out;
        uint64_t factor = 2 * (uint64_t)ntt_size;
        uint64_t value = (uint64_t(1) << bit_size);
        if (value <= factor)
            throw std::logic_error("failed to find enough qualifying primes");
        value = value - factor + 1;
        uint64_t lower = uint64_t(1) << (bit_size - 1);
        while (count > 0 && value > lower)
        {
            if (is_prime_u64(value))
            {
                out.push_back(value);
                count--;
            }
            value -= factor;
        }
        if (count > 0)
            throw std::logic_error("failed to find enough qualifying primes");
        return out;
    }

    // CoeffModulus::Create (modulus.cpp:143-182): per bit size take primes from the back of
    // the descending list, in the order the sizes are requested.
    inline std::vector<uint64_t> coeff_modulus_create(int log_n, const std::vector<int> &bits)
    {
        if (log_n < 1 || log_n > 17)
            throw std::invalid_argument("poly_modulus_degree is invalid");
        std::unordered_map<int, size_t> cnt;
        for (int b : bits)
            cnt[b]++;
        std::unordered_map<int, std::vector<uint64_t>> table;
        for (auto &kv : cnt)
            table[kv.first] = get_primes(size_t(1) << log_n, kv.first, kv.second);
        std::vector<uint64_t> out;
        for (int b : bits)
        {
            out.push_back(table[b].back());
            table[b].pop_back();
        }
        return out;
    }

    // try_minimal_primitive_root (util/numth.cpp:398-425).  SEAL starts from a random primitive
    // root and scans all odd powers for the minimum; the minimum is independent of the start, so
    // we start from a deterministic one.
    inline uint64_t minimal_primitive_root(uint64_t degree /* 2N */, uint64_t q)
    {
        if ((q - 1) % degree != 0)
            throw std::invalid_argument("invalid modulus");
        uint64_t quot = (q - 1) / degree;
        uint64_t root = 0;
        for (uint64_t g = 2; g < 1000; g++)
        {
            uint64_t r = powmod(g, quot, q);
            if (powmod(r, degree >> 1, q) == q - 1)
            {
                root = r;
                break;
            }
        }
        if (!root)
            throw std::invalid_argument("invalid modulus");
        uint64_t gen_sq = mulmod(root, root, q);
        uint64_t cur = root;
        uint64_t best = root;
        for (uint64_t i = 0; i < degree / 2; i++)
        {
            if (cur < best)
                best = cur;
            cur = mulmod(cur, gen_sq, q);
        }
        return best;
    }

    // GaloisTool::get_elt_from_step (util/galois.cpp:53-95), generator 5 (galois.h:169).
    inline uint32_t galois_elt_from_step(int log_n, int step)
    {
        uint32_t n = uint32_t(1) << log_n;
        uint64_t m = 2ull * n;
        if (step == 0)
            return (uint32_t)(m - 1);
        bool neg = step < 0;
        uint32_t pos = (uint32_t)(neg ? -(int64_t)step : step);
        if (pos >= (n >> 1))
            throw std::invalid_argument("step count too large");
        uint32_t s = neg ? (n >> 1) - pos : pos;
        uint64_t elt = 1;
        while (s--)
        {
            elt *= 5;
            elt &= m - 1;
        }
        return (uint32_t)elt;
    }

    // GaloisTool::generate_table_ntt (util/galois.cpp:18-51): result[i] = operand[table[i]].
    inline void galois_table_ntt(int log_n, uint32_t elt, uint32_t *table)
    {
        uint32_t n = uint32_t(1) << log_n;
        if (!(elt & 1) || elt >= 2 * (uint64_t)n)
            throw std::invalid_argument("Galois element is not valid");
        for (uint32_t i = 0; i < n; i++)
        {
            uint32_t rev = bitrev(n + i, log_n + 1);
            uint64_t raw = ((uint64_t)elt * rev) >> 1;
            raw &= (uint64_t)(n - 1);
            table[i] = bitrev((uint32_t)raw, log_n);
        }
    }

    // NTTTables::initialize (util/ntt.cpp:30-89).  tw[k] = psi^bitrev(k) (forward, SEAL's
    // root_powers_ order).  For the inverse we keep the SAME indexing, itw[k] = tw[k]^-1, which
    // is the exact inverse butterfly of stage/group k; SEAL stores the same values in a
    // "scrambled" order (ntt.cpp:69-77) - see seal_inv_root_powers() for that order.
    inline void ntt_tables(int log_n, uint64_t q, std::vector<uint64_t> &tw, std::vector<uint64_t> &itw)
    {
        size_t n = size_t(1) << log_n;
        uint64_t psi = minimal_primitive_root(2 * n, q);
        uint64_t ipsi = invmod(psi, q);
        tw.assign(n, 0);
        itw.assign(n, 0);
        uint64_t p = 1, ip = 1;
        for (size_t i = 0; i < n; i++)
        {
            uint32_t r = bitrev((uint32_t)i, log_n);
            tw[r] = p;
            itw[r] = ip;
            p = mulmod(p, psi, q);
            ip = mulmod(ip, ipsi, q);
        }
    }

    // inv_root_powers_ in SEAL's own order (ntt.cpp:69-77): [bitrev(i-1)+1] = psi^-i, [0] = 1.
    inline std::vector<uint64_t> seal_inv_root_powers(int log_n, uint64_t q)
    {
        size_t n = size_t(1) << log_n;
        uint64_t ipsi = invmod(minimal_primitive_root(2 * n, q), q);
        std::vector<uint64_t> out(n);
        uint64_t p = ipsi;
        for (size_t i = 1; i < n; i++)
        {
            out[bitrev((uint32_t)(i - 1), log_n) + 1] = p;
            p = mulmod(p, ipsi, q);
        }
        out[0] = 1;
        return out;
    }
} // namespace bk
