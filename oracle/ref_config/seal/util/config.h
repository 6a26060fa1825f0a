// Hand-written build configuration for compiling the reference's modified
// SEAL 3.6.6 *in place* (sources stay under /root/reference) into oracle/_ref/.
// Replaces the header that the reference's CMake would generate from
// seal/util/config.h.in; third-party options (MSGSL, zlib, zstd, HEXL) are off.
// TEST INFRASTRUCTURE ONLY - nothing in the product path includes this.
#pragma once

#define SEAL_VERSION "3.6.6"
#define SEAL_VERSION_MAJOR 3
#define SEAL_VERSION_MINOR 6
#define SEAL_VERSION_PATCH 6

#define SEAL_USE_STD_BYTE
#define SEAL_USE_ALIGNED_ALLOC
#define SEAL_USE_SHARED_MUTEX
#define SEAL_USE_IF_CONSTEXPR
#define SEAL_USE_MAYBE_UNUSED
#define SEAL_USE_NODISCARD
#define SEAL_USE_STD_FOR_EACH_N

#define SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT
#define SEAL_DEFAULT_PRNG Blake2xb

#define SEAL_USE_INTRIN
#define SEAL_USE___BUILTIN_CLZLL
#define SEAL_USE___INT128
#define SEAL_USE__ADDCARRY_U64
#define SEAL_USE__SUBBORROW_U64

#define SEAL_USE_EXPLICIT_BZERO
