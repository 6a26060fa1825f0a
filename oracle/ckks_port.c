/* oracle/ckks_port.c - TEST INFRASTRUCTURE ONLY (the checker; never linked, imported or
 * executed by the product path under fhe-gpt-2_b200/).
 *
 * Plain-C restatement of the reference's hot-path ALGORITHMS (modified SEAL 3.6.6 under
 * cnn_ckks/cpu-ckks/single-key/seal-modified-3.6.6/native/src/seal/), each function citing
 * the file:line it follows.  Parity status: PINNED - tests/test_oracle_cpu.py checks this file
 * against (a) the known-answer vectors of the reference's own unit tests
 * (native/tests/seal/util/{ntt,rns,galois,uintarithsmallmod,polyarithsmallmod}.cpp) and
 * (b) outputs of the reference itself (oracle/_ref/libseal_ref.so, compiled in place from
 * /root/reference) on seeded inputs, incl. a full key switch and rescale at N = 4096..65536.
 *
 * Scalar, one thread, unsigned __int128 for the wide products: written for obviousness, not
 * speed.  Layouts: limb arrays [limb][coeff]; ciphertext [poly][limb][coeff]; key-switch key
 * [digit][poly][key limb][coeff] (kswitchkeys.h:340).
 */
#include "ckks_port.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---- util/uintarithsmallmod.h -------------------------------------------------------------- */

/* barrett_reduce_128 (uintarithsmallmod.h:167-204): input < 2^128 -> input mod q */
uint64_t port_barrett_reduce_128(uint64_t lo, uint64_t hi, uint64_t q)
{
    /* const_ratio = floor(2^128 / q) (modulus.cpp set_value) */
    u128 ratio = (~(u128)0) / q; /* == floor(2^128/q) for q not a power of two */
    if ((q & (q - 1)) == 0)
        ratio = q == 1 ? ~(u128)0 : ((u128)1 << 127) / (q >> 1);
    uint64_t r0 = (uint64_t)ratio, r1 = (uint64_t)(ratio >> 64);
    /* round 1 */
    uint64_t carry = (uint64_t)(((u128)lo * r0) >> 64);
    u128 t2 = (u128)lo * r1;
    uint64_t tmp1 = (uint64_t)t2 + carry;
    uint64_t tmp3 = (uint64_t)(t2 >> 64) + (tmp1 < (uint64_t)t2);
    /* round 2 */
    t2 = (u128)hi * r0;
    uint64_t s = tmp1 + (uint64_t)t2;
    carry = (uint64_t)(t2 >> 64) + (s < tmp1);
    uint64_t quot = hi * r1 + tmp3 + carry;
    uint64_t r = lo - quot * q;
    return r >= q ? r - q : r;
}

/* barrett_reduce_64 (uintarithsmallmod.h:211-240) */
uint64_t port_barrett_reduce_64(uint64_t x, uint64_t q)
{
    u128 ratio = (~(u128)0) / q;
    if ((q & (q - 1)) == 0)
        ratio = q == 1 ? ~(u128)0 : ((u128)1 << 127) / (q >> 1);
    uint64_t r1 = (uint64_t)(ratio >> 64);
    uint64_t t = (uint64_t)(((u128)x * r1) >> 64);
    uint64_t r = x - t * q;
    return r >= q ? r - q : r;
}

/* multiply_uint_mod (uintarithsmallmod.h:242-253) */
uint64_t port_mulmod(uint64_t a, uint64_t b, uint64_t q)
{
    u128 p = (u128)a * b;
    return port_barrett_reduce_128((uint64_t)p, (uint64_t)(p >> 64), q);
}

/* MultiplyUIntModOperand::set_quotient (uintarithsmallmod.h:255-285): floor(w * 2^64 / q) */
uint64_t port_shoup_quotient(uint64_t w, uint64_t q)
{
    return (uint64_t)(((u128)w << 64) / q);
}

/* multiply_uint_mod_lazy (uintarithsmallmod.h:313-326): result in [0, 2q) */
static inline uint64_t mul_lazy(uint64_t x, uint64_t w, uint64_t wq, uint64_t q)
{
    uint64_t hi = (uint64_t)(((u128)x * wq) >> 64);
    return x * w - hi * q;
}

/* multiply_uint_mod with operand (uintarithsmallmod.h:290-311): result in [0, q) */
uint64_t port_mulmod_operand(uint64_t x, uint64_t w, uint64_t wq, uint64_t q)
{
    uint64_t r = mul_lazy(x, w, wq, q);
    return r >= q ? r - q : r;
}

static uint64_t powmod(uint64_t a, uint64_t e, uint64_t q)
{
    uint64_t r = 1 % q;
    a %= q;
    while (e)
    {
        if (e & 1)
            r = port_mulmod(r, a, q);
        a = port_mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}

static uint32_t bitrev32(uint32_t x, int bits)
{
    uint32_t r = 0;
    for (int i = 0; i < bits; i++)
    {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return bits ? r : 0;
}

/* ---- util/numth.cpp ------------------------------------------------------------------------ */

/* try_minimal_primitive_root (numth.cpp:398-425): the smallest primitive degree-th root.
 * SEAL finds one primitive root at random (try_primitive_root :356-396) and then walks all
 * odd powers keeping the minimum; the minimum does not depend on the starting root. */
uint64_t port_minimal_primitive_root(uint64_t degree, uint64_t q)
{
    uint64_t size_quotient_group = (q - 1) / degree;
    if (q - 1 != size_quotient_group * degree)
        return 0;
    uint64_t root = 0;
    for (uint64_t g = 2; g < 100000 && !root; g++)
    {
        uint64_t r = powmod(g, size_quotient_group, q);
        /* is_primitive_root (numth.cpp:341-354): r^(degree/2) == -1 */
        if (powmod(r, degree >> 1, q) == q - 1)
            root = r;
    }
    if (!root)
        return 0;
    uint64_t generator_sq = port_mulmod(root, root, q);
    uint64_t current = root, best = root;
    for (uint64_t i = 0; i < degree; i += 2)
    {
        if (current < best)
            best = current;
        current = port_mulmod(current, generator_sq, q);
    }
    return best;
}

/* ---- util/ntt.cpp:30-89  NTTTables::initialize ------------------------------------------- */
int port_ntt_tables_init(port_ntt_tables *t, int log_n, uint64_t q)
{
    size_t n = (size_t)1 << log_n;
    memset(t, 0, sizeof(*t));
    t->log_n = log_n;
    t->n = n;
    t->q = q;
    t->root = port_minimal_primitive_root(2 * n, q);
    if (!t->root)
        return -1;
    t->root_powers = malloc(n * sizeof(uint64_t));
    t->root_powers_q = malloc(n * sizeof(uint64_t));
    t->inv_root_powers = malloc(n * sizeof(uint64_t));
    t->inv_root_powers_q = malloc(n * sizeof(uint64_t));
    uint64_t inv_root = powmod(t->root, q - 2, q);
    /* root_powers_[reverse_bits(i)] = root^i  (ntt.cpp:58-67) */
    uint64_t p = 1;
    t->root_powers[0] = 1;
    for (size_t i = 1; i < n; i++)
    {
        p = port_mulmod(p, t->root, q);
        t->root_powers[bitrev32((uint32_t)i, log_n)] = p;
    }
    /* inv_root_powers_[reverse_bits(i - 1) + 1] = inv_root^i  (ntt.cpp:69-77) */
    p = 1;
    t->inv_root_powers[0] = 1;
    for (size_t i = 1; i < n; i++)
    {
        p = port_mulmod(p, inv_root, q);
        t->inv_root_powers[bitrev32((uint32_t)(i - 1), log_n) + 1] = p;
    }
    for (size_t i = 0; i < n; i++)
    {
        t->root_powers_q[i] = port_shoup_quotient(t->root_powers[i], q);
        t->inv_root_powers_q[i] = port_shoup_quotient(t->inv_root_powers[i], q);
    }
    /* inv_degree_modulo_ (ntt.cpp:79-84) */
    t->inv_n = powmod((uint64_t)n % q, q - 2, q);
    t->inv_n_q = port_shoup_quotient(t->inv_n, q);
    return 0;
}

void port_ntt_tables_free(port_ntt_tables *t)
{
    free(t->root_powers);
    free(t->root_powers_q);
    free(t->inv_root_powers);
    free(t->inv_root_powers_q);
    memset(t, 0, sizeof(*t));
}

/* ---- util/dwthandler.h:94-191 transform_to_rev with the modular Arithmetic of ntt.h:24-71 -- */
/* lazy: outputs in [0, 4q) (ntt_negacyclic_harvey_lazy, ntt.cpp:183-195) */
void port_ntt_lazy(uint64_t *v, const port_ntt_tables *t)
{
    const uint64_t q = t->q, two_q = 2 * q;
    size_t n = t->n, gap = n >> 1, m = 1, root_idx = 0;
    for (; m < n; m <<= 1, gap >>= 1)
    {
        size_t offset = 0;
        for (size_t i = 0; i < m; i++)
        {
            ++root_idx;
            uint64_t w = t->root_powers[root_idx], wq = t->root_powers_q[root_idx];
            uint64_t *x = v + offset, *y = x + gap;
            for (size_t j = 0; j < gap; j++)
            {
                /* guard (ntt.h:62-65), mul_root (ntt.h:45-49), add / sub (ntt.h:35-43) */
                uint64_t u = x[j] >= two_q ? x[j] - two_q : x[j];
                uint64_t w_y = mul_lazy(y[j], w, wq, q);
                x[j] = u + w_y;
                y[j] = u + two_q - w_y;
            }
            offset += gap << 1;
        }
    }
}

/* ntt_negacyclic_harvey (ntt.h:235-264): lazy transform + correction to [0, q) */
void port_ntt(uint64_t *v, const port_ntt_tables *t)
{
    port_ntt_lazy(v, t);
    const uint64_t q = t->q, two_q = 2 * q;
    for (size_t i = 0; i < t->n; i++)
    {
        if (v[i] >= two_q)
            v[i] -= two_q;
        if (v[i] >= q)
            v[i] -= q;
    }
}

/* dwthandler.h:202-356 transform_from_rev with scalar = n^-1; outputs in [0, 2q)
 * (inverse_ntt_negacyclic_harvey_lazy, ntt.cpp:197-209) */
void port_intt_lazy(uint64_t *v, const port_ntt_tables *t)
{
    const uint64_t q = t->q, two_q = 2 * q;
    size_t n = t->n, gap = 1, m = n >> 1, root_idx = 0;
    for (; m > 1; m >>= 1, gap <<= 1)
    {
        size_t offset = 0;
        for (size_t i = 0; i < m; i++)
        {
            ++root_idx;
            uint64_t w = t->inv_root_powers[root_idx], wq = t->inv_root_powers_q[root_idx];
            uint64_t *x = v + offset, *y = x + gap;
            for (size_t j = 0; j < gap; j++)
            {
                uint64_t u = x[j], vv = y[j];
                uint64_t s = u + vv;
                x[j] = s >= two_q ? s - two_q : s;
                y[j] = mul_lazy(u + two_q - vv, w, wq, q);
            }
            offset += gap << 1;
        }
    }
    /* last stage with the scalar folded in (dwthandler.h:273-314) */
    ++root_idx;
    uint64_t r = t->inv_root_powers[root_idx];
    uint64_t scaled_r = port_mulmod(r, t->inv_n, q); /* mul_root_scalar (ntt.h:56-60) */
    uint64_t scaled_r_q = port_shoup_quotient(scaled_r, q);
    uint64_t *x = v, *y = v + gap;
    for (size_t j = 0; j < gap; j++)
    {
        uint64_t u = x[j] >= two_q ? x[j] - two_q : x[j];
        uint64_t vv = y[j];
        uint64_t s = u + vv;
        s = s >= two_q ? s - two_q : s;
        x[j] = mul_lazy(s, t->inv_n, t->inv_n_q, q);
        y[j] = mul_lazy(u + two_q - vv, scaled_r, scaled_r_q, q);
    }
}

/* inverse_ntt_negacyclic_harvey (ntt.h:336-358) */
void port_intt(uint64_t *v, const port_ntt_tables *t)
{
    port_intt_lazy(v, t);
    for (size_t i = 0; i < t->n; i++)
        if (v[i] >= t->q)
            v[i] -= t->q;
}

/* ---- util/polyarithsmallmod.cpp ------------------------------------------------------------- */
/* dyadic_product_coeffmod (:111-169) */
void port_dyadic_product(const uint64_t *a, const uint64_t *b, size_t n, uint64_t q, uint64_t *out)
{
    for (size_t i = 0; i < n; i++)
        out[i] = port_mulmod(a[i], b[i], q);
}
/* add_poly_coeffmod / sub_poly_coeffmod / negate_poly_coeffmod (:18-80) */
void port_add_poly(const uint64_t *a, const uint64_t *b, size_t n, uint64_t q, uint64_t *out)
{
    for (size_t i = 0; i < n; i++)
    {
        uint64_t s = a[i] + b[i];
        out[i] = s >= q ? s - q : s;
    }
}
void port_sub_poly(const uint64_t *a, const uint64_t *b, size_t n, uint64_t q, uint64_t *out)
{
    for (size_t i = 0; i < n; i++)
        out[i] = a[i] >= b[i] ? a[i] - b[i] : a[i] + q - b[i];
}
void port_negate_poly(const uint64_t *a, size_t n, uint64_t q, uint64_t *out)
{
    for (size_t i = 0; i < n; i++)
        out[i] = a[i] ? q - a[i] : 0;
}

/* ---- util/galois.cpp ------------------------------------------------------------------------ */
/* get_elt_from_step (:53-95), generator 5 in this fork (galois.h:169) */
uint32_t port_galois_elt_from_step(int log_n, int step)
{
    uint32_t n = (uint32_t)1 << log_n;
    uint32_t m32 = n << 1;
    uint64_t m = m32;
    if (step == 0)
        return (uint32_t)(m - 1);
    int neg = step < 0;
    uint32_t pos_step = (uint32_t)(neg ? -step : step);
    if (pos_step >= (n >> 1))
        return 0;
    pos_step &= m32 - 1;
    uint32_t s = neg ? (n >> 1) - pos_step : pos_step;
    uint64_t gen = 5, elt = 1;
    for (uint32_t i = 0; i < s; i++)
    {
        elt *= gen;
        elt &= m - 1;
    }
    return (uint32_t)elt;
}

/* generate_table_ntt (:18-51) */
void port_galois_table_ntt(int log_n, uint32_t elt, uint32_t *table)
{
    uint32_t n = (uint32_t)1 << log_n;
    uint32_t mask = n - 1;
    for (uint32_t i = 0; i < n; i++)
    {
        uint32_t reversed = bitrev32(n + i, log_n + 1);
        uint64_t index_raw = ((uint64_t)elt * reversed) >> 1;
        index_raw &= mask;
        table[i] = bitrev32((uint32_t)index_raw, log_n);
    }
}

/* apply_galois_ntt (:192-218): result[i] = operand[table[i]] */
void port_apply_galois_ntt(const uint64_t *in, int log_n, uint32_t elt, uint64_t *out)
{
    size_t n = (size_t)1 << log_n;
    uint32_t *table = malloc(n * sizeof(uint32_t));
    port_galois_table_ntt(log_n, elt, table);
    for (size_t i = 0; i < n; i++)
        out[i] = in[table[i]];
    free(table);
}

/* apply_galois (:97-160), coefficient form: X^i -> X^(i*elt), sign flip on wrap-around */
void port_apply_galois(const uint64_t *in, int log_n, uint32_t elt, uint64_t q, uint64_t *out)
{
    size_t n = (size_t)1 << log_n;
    uint64_t index_raw = 0;
    for (size_t i = 0; i < n; i++, index_raw += elt)
    {
        uint64_t index = index_raw & (n - 1);
        uint64_t v = in[i];
        if ((index_raw >> log_n) & 1)
            v = v ? q - v : 0;
        out[index] = v;
    }
}

/* ---- util/rns.cpp:737-808  divide_and_round_q_last_ntt_inplace (the rescale kernel) ---- */
/* poly: [limbs][n] NTT form; tables[i] for limb i.  On return limbs 0..limbs-2 hold the result. */
void port_divide_and_round_q_last_ntt(uint64_t *poly, int limbs, const port_ntt_tables *tables)
{
    size_t n = tables[0].n;
    uint64_t *last = poly + (size_t)(limbs - 1) * n;
    uint64_t qk = tables[limbs - 1].q;
    port_intt(last, &tables[limbs - 1]);
    uint64_t half = qk >> 1;
    for (size_t i = 0; i < n; i++)
    {
        uint64_t s = last[i] + half; /* add_poly_scalar_coeffmod */
        last[i] = s >= qk ? s - qk : s;
    }
    uint64_t *temp = malloc(n * sizeof(uint64_t));
    for (int l = 0; l < limbs - 1; l++)
    {
        uint64_t qi = tables[l].q;
        for (size_t i = 0; i < n; i++)
            temp[i] = qi < qk ? port_barrett_reduce_64(last[i], qi) : last[i];
        uint64_t neg_half_mod = qi - port_barrett_reduce_64(half, qi);
        for (size_t i = 0; i < n; i++)
            temp[i] += neg_half_mod;
        uint64_t qi_lazy = qi << 2;
        port_ntt_lazy(temp, &tables[l]);
        uint64_t inv = powmod(qk % qi, qi - 2, qi); /* inv_q_last_mod_q_ (rns.cpp:300-312) */
        uint64_t inv_q = port_shoup_quotient(inv, qi);
        uint64_t *x = poly + (size_t)l * n;
        for (size_t i = 0; i < n; i++)
        {
            x[i] += qi_lazy - temp[i];
            /* multiply_poly_scalar_coeffmod (polyarithsmallmod.cpp:82-109) */
            x[i] = port_mulmod_operand(x[i], inv, inv_q, qi);
        }
    }
    free(temp);
}

/* ---- evaluator.cpp:2281-2525  switch_key_inplace (CKKS branch) ------------------------------ */
/* ct: [2][l][n] (updated in place); target: [l][n] NTT form; key: [digits][2][key_limbs][n];
 * tables: one per key-level prime (key_limbs of them, special prime last). */
void port_switch_key(uint64_t *ct, const uint64_t *target, const uint64_t *key, int l, int key_limbs,
                     const port_ntt_tables *tables)
{
    size_t n = tables[0].n;
    int rns = l + 1;
    uint64_t *t_target = malloc((size_t)l * n * sizeof(uint64_t));
    memcpy(t_target, target, (size_t)l * n * sizeof(uint64_t));
    for (int j = 0; j < l; j++)
        port_intt(t_target + (size_t)j * n, &tables[j]); /* :2358-2365 */
    uint64_t *t_poly_prod = calloc((size_t)2 * rns * n, sizeof(uint64_t));
    uint64_t *t_ntt = malloc(n * sizeof(uint64_t));
    u128 *lazy = malloc((size_t)2 * n * sizeof(u128));
    for (int I = 0; I < rns; I++)
    {
        int key_index = I == l ? key_limbs - 1 : I; /* :2371 */
        uint64_t qI = tables[key_index].q;
        memset(lazy, 0, (size_t)2 * n * sizeof(u128));
        for (int J = 0; J < l; J++)
        {
            const uint64_t *operand;
            if (I == J)
                operand = target + (size_t)J * n; /* :2389-2392 */
            else
            {
                uint64_t qJ = tables[J].q;
                for (size_t i = 0; i < n; i++) /* :2397-2405 */
                    t_ntt[i] = qJ <= qI ? t_target[(size_t)J * n + i] : port_barrett_reduce_64(t_target[(size_t)J * n + i], qI);
                port_ntt_lazy(t_ntt, &tables[key_index]); /* :2407 */
                operand = t_ntt;
            }
            for (int K = 0; K < 2; K++) /* :2412-2436, lazy 128-bit accumulation (never > 256 summands here) */
            {
                const uint64_t *k = key + (((size_t)J * 2 + K) * key_limbs + key_index) * n;
                for (size_t i = 0; i < n; i++)
                    lazy[(size_t)K * n + i] += (u128)operand[i] * k[i];
            }
        }
        for (int K = 0; K < 2; K++) /* :2448-2462 */
            for (size_t i = 0; i < n; i++)
            {
                u128 a = lazy[(size_t)K * n + i];
                t_poly_prod[((size_t)K * rns + I) * n + i] = port_barrett_reduce_128((uint64_t)a, (uint64_t)(a >> 64), qI);
            }
    }
    /* modulus switching with scaling (:2465-2523) */
    uint64_t qk = tables[key_limbs - 1].q, qk_half = qk >> 1;
    for (int K = 0; K < 2; K++)
    {
        uint64_t *t_last = t_poly_prod + ((size_t)K * rns + l) * n;
        port_intt_lazy(t_last, &tables[key_limbs - 1]);
        for (size_t i = 0; i < n; i++)
            t_last[i] = port_barrett_reduce_64(t_last[i] + qk_half, qk);
        for (int J = 0; J < l; J++)
        {
            uint64_t qi = tables[J].q;
            for (size_t i = 0; i < n; i++)
                t_ntt[i] = qk > qi ? port_barrett_reduce_64(t_last[i], qi) : t_last[i];
            uint64_t fix = qi - port_barrett_reduce_64(qk_half, qi);
            for (size_t i = 0; i < n; i++)
                t_ntt[i] += fix;
            uint64_t qi_lazy = qi << 2;
            port_ntt_lazy(t_ntt, &tables[J]);
            uint64_t inv = powmod(qk % qi, qi - 2, qi); /* key-level inv_q_last_mod_q (:2336) */
            uint64_t inv_q = port_shoup_quotient(inv, qi);
            uint64_t *prod = t_poly_prod + ((size_t)K * rns + J) * n;
            uint64_t *dst = ct + ((size_t)K * l + J) * n;
            for (size_t i = 0; i < n; i++)
            {
                prod[i] += qi_lazy - t_ntt[i];
                prod[i] = port_mulmod_operand(prod[i], inv, inv_q, qi);
                uint64_t s = prod[i] + dst[i]; /* add_poly_coeffmod */
                dst[i] = s >= qi ? s - qi : s;
            }
        }
    }
    free(lazy);
    free(t_ntt);
    free(t_poly_prod);
    free(t_target);
}

/* ---- evaluator.cpp:2120-2222  apply_galois_inplace (CKKS): permute c0, permute c1 into the
 * key-switch target, zero c1, switch key ------------------------------------------------------ */
void port_apply_galois_ct(uint64_t *ct, int l, uint32_t elt, const uint64_t *key, int key_limbs, int log_n,
                          const port_ntt_tables *tables)
{
    size_t n = (size_t)1 << log_n;
    uint64_t *temp = malloc((size_t)l * n * sizeof(uint64_t));
    for (int j = 0; j < l; j++)
        port_apply_galois_ntt(ct + (size_t)j * n, log_n, elt, temp + (size_t)j * n);
    memcpy(ct, temp, (size_t)l * n * sizeof(uint64_t));
    for (int j = 0; j < l; j++)
        port_apply_galois_ntt(ct + ((size_t)l + j) * n, log_n, elt, temp + (size_t)j * n);
    memset(ct + (size_t)l * n, 0, (size_t)l * n * sizeof(uint64_t));
    port_switch_key(ct, temp, key, l, key_limbs, tables);
    free(temp);
}

/* ---- evaluator.cpp:744-772  ckks_multiply, size 2 x size 2 ------------------------------------ */
/* a, b: [2][l][n]; out: [3][l][n] */
void port_ckks_multiply(const uint64_t *a, const uint64_t *b, uint64_t *out, int l, size_t n, const uint64_t *primes)
{
    for (int j = 0; j < l; j++)
    {
        uint64_t q = primes[j];
        const uint64_t *a0 = a + (size_t)j * n, *a1 = a + ((size_t)l + j) * n;
        const uint64_t *b0 = b + (size_t)j * n, *b1 = b + ((size_t)l + j) * n;
        for (size_t i = 0; i < n; i++)
        {
            out[(size_t)j * n + i] = port_mulmod(a0[i], b0[i], q);
            uint64_t x = port_mulmod(a0[i], b1[i], q), y = port_mulmod(a1[i], b0[i], q);
            uint64_t s = x + y;
            out[((size_t)l + j) * n + i] = s >= q ? s - q : s;
            out[((size_t)2 * l + j) * n + i] = port_mulmod(a1[i], b1[i], q);
        }
    }
}

/* ---- ckks_bootstrapping/Bootstrapper.cpp:2894-2948  modraise (coefficient-form part) ------ */
/* src: [n] residues mod q0 (coefficient form); dst: [limbs][n] */
void port_modraise_coeffs(const uint64_t *src, uint64_t *dst, int limbs, size_t n, const uint64_t *primes)
{
    uint64_t q0 = primes[0];
    for (int j = 0; j < limbs; j++)
    {
        uint64_t q = primes[j];
        uint64_t minus_q0 = j == 0 ? 0 : q - q0 % q;
        for (size_t i = 0; i < n; i++)
        {
            uint64_t v = src[i] % q;
            if (src[i] > (q0 >> 1))
            {
                v += minus_q0;
                v -= v >= q ? q : 0;
            }
            dst[(size_t)j * n + i] = v;
        }
    }
}
