/* oracle/ckks_port.h - TEST INFRASTRUCTURE ONLY. See ckks_port.c. */
#ifndef CKKS_PORT_H
#define CKKS_PORT_H
#include <stddef.h>
#include <stdint.h>

typedef struct
{
    int log_n;
    size_t n;
    uint64_t q, root;
    uint64_t *root_powers, *root_powers_q;         /* ntt.cpp:58-67 order */
    uint64_t *inv_root_powers, *inv_root_powers_q; /* ntt.cpp:69-77 order */
    uint64_t inv_n, inv_n_q;
} port_ntt_tables;

uint64_t port_barrett_reduce_128(uint64_t lo, uint64_t hi, uint64_t q);
uint64_t port_barrett_reduce_64(uint64_t x, uint64_t q);
uint64_t port_mulmod(uint64_t a, uint64_t b, uint64_t q);
uint64_t port_shoup_quotient(uint64_t w, uint64_t q);
uint64_t port_mulmod_operand(uint64_t x, uint64_t w, uint64_t wq, uint64_t q);
uint64_t port_minimal_primitive_root(uint64_t degree, uint64_t q);
int port_ntt_tables_init(port_ntt_tables *t, int log_n, uint64_t q);
void port_ntt_tables_free(port_ntt_tables *t);
void port_ntt_lazy(uint64_t *v, const port_ntt_tables *t);
void port_ntt(uint64_t *v, const port_ntt_tables *t);
void port_intt_lazy(uint64_t *v, const port_ntt_tables *t);
void port_intt(uint64_t *v, const port_ntt_tables *t);
void port_dyadic_product(const uint64_t *a, const uint64_t *b, size_t n, uint64_t q, uint64_t *out);
void port_add_poly(const uint64_t *a, const uint64_t *b, size_t n, uint64_t q, uint64_t *out);
void port_sub_poly(const uint64_t *a, const uint64_t *b, size_t n, uint64_t q, uint64_t *out);
void port_negate_poly(const uint64_t *a, size_t n, uint64_t q, uint64_t *out);
uint32_t port_galois_elt_from_step(int log_n, int step);
void port_galois_table_ntt(int log_n, uint32_t elt, uint32_t *table);
void port_apply_galois_ntt(const uint64_t *in, int log_n, uint32_t elt, uint64_t *out);
void port_apply_galois(const uint64_t *in, int log_n, uint32_t elt, uint64_t q, uint64_t *out);
void port_divide_and_round_q_last_ntt(uint64_t *poly, int limbs, const port_ntt_tables *tables);
void port_switch_key(uint64_t *ct, const uint64_t *target, const uint64_t *key, int l, int key_limbs,
                     const port_ntt_tables *tables);
void port_apply_galois_ct(uint64_t *ct, int l, uint32_t elt, const uint64_t *key, int key_limbs, int log_n,
                          const port_ntt_tables *tables);
void port_ckks_multiply(const uint64_t *a, const uint64_t *b, uint64_t *out, int l, size_t n, const uint64_t *primes);
void port_modraise_coeffs(const uint64_t *src, uint64_t *dst, int limbs, size_t n, const uint64_t *primes);
#endif
