// cnn/cnn_seal.cpp - see cnn_seal.h.  The host side only builds plaintext slot vectors (weights, masks) and decides
// the order of Evaluator calls; rotations, plaintext multiplications, rescales and additions all run on the GPU.
#include "cnn/cnn_seal.h"
#include "common/cached.h"
#include "common/func.h"
#include <cmath>
#include <stdexcept>

using namespace seal;
using minicomp::floor_to_int;
using minicomp::log2_long;
using minicomp::pow2;
using std::vector;

TensorCipher::TensorCipher(int logn, int k, int h, int w, int c, int t, int p, vector<double> data, Encryptor &encryptor,
                           CKKSEncoder &encoder, int logp)
    : k_(k), h_(h), w_(w), c_(c), t_(t), p_(p), logn_(logn)
{
    if (k != 1)
        throw std::invalid_argument("supported k is only 1 right now");
    if (logn < 1 || logn > 16)
        throw std::out_of_range("the value of logn is out of range");
    if (data.size() > (std::size_t(1) << logn))
        throw std::out_of_range("the size of data is larger than n");
    data.resize(std::size_t(1) << logn, 0.0);
    Plaintext plain;
    encoder.encode(data, std::pow(2.0, logp), plain);
    encryptor.encrypt(plain, cipher_);
}

TensorCipher::TensorCipher(int logn, int k, int h, int w, int c, int t, int p, Ciphertext cipher)
    : k_(k), h_(h), w_(w), c_(c), t_(t), p_(p), logn_(logn), cipher_(std::move(cipher))
{}

namespace
{
    // sum += term, or sum = term for the first one
    struct Accumulator
    {
        Evaluator &ev;
        Ciphertext &sum;
        bool started = false;
        void add(const Ciphertext &term)
        {
            if (!started)
            {
                sum = term;
                started = true;
            }
            else
                ev.add_inplace_reduced_error(sum, term);
        }
    };

    void rotated_copy(const Ciphertext &src, Ciphertext &dst, int steps, Evaluator &evaluator, GaloisKeys &gal_keys)
    {
        const long n = (long)src.poly_modulus_degree() / 2;
        if (&dst == &src || ((steps % n) + n) % n == 0)
        {
            dst = src;
            memory_save_rotate(dst, dst, steps, evaluator, gal_keys);
        }
        else
            memory_save_rotate(src, dst, steps, evaluator, gal_keys); // rotates straight into dst
    }

    // dst[i] = src rotated by steps[i].  The reference rotates one by one (rotated_copy); on the engine the rotations
    // of ONE ciphertext share the decomposition of its c1 (Evaluator::rotate_vector_hoisted; tolerance mode, off with
    // $B200CKKS_NO_HOIST or the reference's constants) - steps that memory_save_rotate splits in two keep that path.
    void rotated_copies(const Ciphertext &src, const vector<int> &steps, vector<Ciphertext> &dst, Evaluator &evaluator,
                        GaloisKeys &gal_keys)
    {
        dst.resize(steps.size());
#ifdef B200CKKS_FACADE
        static const bool hoist = std::getenv("B200CKKS_NO_HOIST") == nullptr && !encrypt_constants();
        if (hoist && steps.size() > 1)
        {
            const long n = (long)src.poly_modulus_degree() / 2;
            vector<int> hs;
            vector<std::size_t> where;
            for (std::size_t i = 0; i < steps.size(); i++)
            {
                const int st = (int)(((steps[i] % n) + n) % n);
                const bool split = (34 <= st && st <= 55) || (57 <= st && st <= 61);
                if (split)
                    rotated_copy(src, dst[i], steps[i], evaluator, gal_keys);
                else
                {
                    hs.push_back(st);
                    where.push_back(i);
                }
            }
            vector<Ciphertext> out;
            evaluator.rotate_vector_hoisted(src, hs, gal_keys, out);
            for (std::size_t k = 0; k < where.size(); k++)
                dst[where[k]] = std::move(out[k]);
            return;
        }
#endif
        for (std::size_t i = 0; i < steps.size(); i++)
            rotated_copy(src, dst[i], steps[i], evaluator, gal_keys);
    }

    int copies_that_fit(long n, long used)
    {
        return (int)pow2(floor_to_int(std::log(static_cast<double>(n) / static_cast<double>(used)) / std::log(2.0)));
    }
} // namespace

void memory_save_rotate(const Ciphertext &cipher_in, Ciphertext &cipher_out, int steps, Evaluator &evaluator,
                        GaloisKeys &gal_keys)
{
    const long n = (long)cipher_in.poly_modulus_degree() / 2;
    steps = (int)(((steps % n) + n) % n);
    if (steps == 0)
        return; // the reference leaves cipher_out untouched here; every caller passes the same object twice
    Ciphertext temp;
    const int first = ((34 <= steps && steps <= 55) || (57 <= steps && steps <= 61)) ? 33 : 0;
    if (first == 0)
        evaluator.rotate_vector(cipher_in, steps, gal_keys, temp);
    else
    {
        evaluator.rotate_vector(cipher_in, first, gal_keys, temp);
        evaluator.rotate_vector_inplace(temp, steps - first, gal_keys);
    }
    cipher_out = std::move(temp);
}

ConvPlan build_conv_plan(const TensorCipher &cnn_in, int co, int st, int fh, int fw, const vector<double> &data,
                         const vector<double> &running_var, const vector<double> &constant_weight, double epsilon)
{
    ConvPlan P;
    const int ki = cnn_in.k(), hi = cnn_in.h(), wi = cnn_in.w(), ci = cnn_in.c(), ti = cnn_in.t(), pi = cnn_in.p(),
              logn = cnn_in.logn();
    if (st != 1 && st != 2)
        throw std::invalid_argument("supported st is only 1 or 2");
    if ((int)data.size() != fh * fw * ci * co)
        throw std::invalid_argument("the size of data vector is not ker x ker x h x h");
    if (log2_long(ki) == -1)
        throw std::invalid_argument("ki is not power of two");
    if ((int)running_var.size() != co || (int)constant_weight.size() != co)
        throw std::invalid_argument("the size of running_var or weight is not correct");
    for (double v : running_var)
        if (v < 1e-16 && v > -1e-16)
            throw std::invalid_argument("the size of running_var is too small. nearly zero.");
    if (fh % 2 == 0 || fw % 2 == 0)
        throw std::invalid_argument("fh and fw should be odd");
    if (st == 2 && (hi % 2 == 1 || wi % 2 == 1))
        throw std::invalid_argument("hi or wi is not even");

    P.ki = ki, P.hi = hi, P.wi = wi, P.ci = ci, P.ti = ti, P.pi = pi, P.logn = logn;
    P.co = co, P.st = st, P.fh = fh, P.fw = fw;
    P.ho = hi / st, P.wo = wi / st, P.ko = ki * st;
    const int ho = P.ho, wo = P.wo, ko = P.ko;
    const long n = 1L << logn;
    P.to = (co + ko * ko - 1) / (ko * ko);
    P.po = copies_that_fit(n, (long)ko * ko * ho * wo * P.to);
    P.q = (co + pi - 1) / pi;
    const int to = P.to;
    const long q = P.q;
    if (n % pi != 0)
        throw std::out_of_range("n is not divisible by pi");
    if (n % P.po != 0)
        throw std::out_of_range("n is not divisible by po");
    if ((long)ki * ki * hi * wi * ti * pi > n)
        throw std::out_of_range("ki^2 hi wi ti pi is larger than n");
    if ((long)ko * ko * ho * wo * to * P.po > n)
        throw std::out_of_range("ko^2 ho wo to po is larger than n");

    // tap (i1, i2), output-channel group g: slot -> weight of (input channel at that slot, output channel = copy + pi g)
    // shifted by the tap, zero where the tap would read outside the image
    const long per_copy = n / pi, plane = (long)ki * ki * hi * wi, row = (long)ki * wi;
    P.tap_weights.assign((std::size_t)(fh * fw) * (std::size_t)q, vector<double>((std::size_t)n, 0.0));
    for (long slot = 0; slot < n; slot++)
    {
        const long r = slot % per_copy, copy = slot / per_copy;
        if (r >= plane * ti)
            continue;
        const long group = r / plane, yy = (r % plane) / row, xx = r % row;
        const long ch = (long)ki * ki * group + ki * (yy % ki) + xx % ki;
        if (ch >= ci)
            continue;
        const long y = yy / ki, x = xx / ki;
        for (int i1 = 0; i1 < fh; i1++)
        {
            const long ys = y - (fh - 1) / 2 + i1;
            if (ys < 0 || ys > hi - 1)
                continue;
            for (int i2 = 0; i2 < fw; i2++)
            {
                const long xs = x - (fw - 1) / 2 + i2;
                if (xs < 0 || xs > wi - 1)
                    continue;
                for (long g = 0; g < q; g++)
                {
                    const long oc = copy + (long)pi * g;
                    if (oc >= co)
                        continue;
                    P.tap_weights[(std::size_t)((i1 * fw + i2) * q + g)][(std::size_t)slot] =
                        data[(std::size_t)((long)fh * fw * ci * oc + (long)fh * fw * ch + fw * i1 + i2)];
                }
            }
        }
    }
    // output channel j: 1/sqrt(var + eps) * bn weight at the slots of channel j in the output layout, else 0
    P.select_one_vec.assign((std::size_t)co, vector<double>((std::size_t)n, 0.0));
    for (int u = 0; u < to; u++)
        for (int v1 = 0; v1 < ko * ho; v1++)
            for (int v2 = 0; v2 < ko * wo; v2++)
            {
                const int j = ko * ko * u + ko * (v1 % ko) + v2 % ko;
                if (j < co)
                    P.select_one_vec[(std::size_t)j][(std::size_t)((long)ko * ko * ho * wo * u + (long)ko * wo * v1 + v2)] =
                        constant_weight[(std::size_t)j] / std::sqrt(running_var[(std::size_t)j] + epsilon);
            }
    return P;
}

void multiplexed_parallel_convolution_planned(const TensorCipher &cnn_in, TensorCipher &cnn_out, const ConvPlan &P,
                                              CKKSEncoder &encoder, Encryptor &encryptor, Evaluator &evaluator,
                                              GaloisKeys &gal_keys, bool end, bool named)
{
    if (cnn_in.k() != P.ki || cnn_in.h() != P.hi || cnn_in.w() != P.wi || cnn_in.c() != P.ci || cnn_in.t() != P.ti ||
        cnn_in.p() != P.pi || cnn_in.logn() != P.logn)
        throw std::invalid_argument("the convolution plan was built for a different input tensor");
    const int ki = P.ki, hi = P.hi, wi = P.wi, ti = P.ti, pi = P.pi, logn = P.logn, co = P.co, fh = P.fh, fw = P.fw;
    const int ko = P.ko, ho = P.ho, wo = P.wo, to = P.to, po = P.po;
    const long n = 1L << logn, q = P.q;
    const void *owner = named ? (const void *)&P : nullptr;

    Ciphertext ctxt_in = cnn_in.cipher(), ct_zero, temp, sum, total_sum, var;

    // the fh x fw shifted copies of the input
    vector<Ciphertext> ctxt_rot;
    {
        vector<int> tap_steps;
        for (int i1 = 0; i1 < fh; i1++)
            for (int i2 = 0; i2 < fw; i2++)
                tap_steps.push_back(ki * ki * wi * (i1 - (fh - 1) / 2) + ki * (i2 - (fw - 1) / 2));
        rotated_copies(ctxt_in, tap_steps, ctxt_rot, evaluator, gal_keys);
    }

    // an encryption of zero at the input's scale: the reference's start value of its running sums (cnn_seal.cpp:433-436).
    // Without it (default) a running sum starts from its first term - same level and scale as zero + first term.
    const bool zero_start = encrypt_constants();
    if (zero_start)
    {
        vector<double> zero((std::size_t)n, 0.0);
        Plaintext plain;
        encoder.encode(zero, ctxt_in.scale(), plain);
        encryptor.encrypt(plain, ct_zero);
    }

    const int d = (int)log2_long(ki), c = (int)log2_long(ti);
    bool total_started = false;
    for (long g = 0; g < q; g++)
    {
        // weighted sum over the filter taps
        vector<const Ciphertext *> taps;
        vector<std::uint64_t> tap_ids;
        for (int t = 0; t < fh * fw; t++)
        {
            taps.push_back(&ctxt_rot[(std::size_t)t]);
            tap_ids.push_back((std::uint64_t)(t * q + g));
        }
        multiply_vector_named_sum(evaluator, sum, taps, owner, tap_ids, 0,
                                  [&](std::uint64_t idx) -> const vector<double> & { return P.tap_weights[(std::size_t)idx]; });
        evaluator.rescale_to_next_inplace(sum);
        var = sum;

        // sum over the input channels: inside a k x k cell, then over the channel groups
        for (int x = 0; x < d; x++)
        {
            rotated_copy(var, temp, (int)pow2(x), evaluator, gal_keys);
            evaluator.add_inplace_reduced_error(var, temp);
        }
        for (int x = 0; x < d; x++)
        {
            rotated_copy(var, temp, (int)pow2(x) * ki * wi, evaluator, gal_keys);
            evaluator.add_inplace_reduced_error(var, temp);
        }
        if (c == -1)
        {
            if (zero_start)
                sum = ct_zero;
            for (int x = 0; x < ti; x++)
            {
                rotated_copy(var, temp, ki * ki * hi * wi * x, evaluator, gal_keys);
                if (x == 0 && !zero_start)
                    sum = temp;
                else
                    evaluator.add_inplace_reduced_error(sum, temp);
            }
            var = sum;
        }
        else
            for (int x = 0; x < c; x++)
            {
                rotated_copy(var, temp, (int)pow2(x) * ki * ki * hi * wi, evaluator, gal_keys);
                evaluator.add_inplace_reduced_error(var, temp);
            }

        // move every output channel of this group to its place in the output layout and mask it
        vector<int> move_steps;
        for (int i8 = 0; i8 < pi && pi * g + i8 < co; i8++)
        {
            const int j4 = (int)(pi * g) + i8;
            move_steps.push_back((int)((n / pi) * (j4 % pi) - j4 % ko - (long)(j4 / (ko * ko)) * ko * ko * ho * wo -
                                       (long)((j4 % (ko * ko)) / ko) * ko * wo));
        }
        vector<Ciphertext> moved;
        rotated_copies(var, move_steps, moved, evaluator, gal_keys);
        for (std::size_t i8 = 0; i8 < move_steps.size(); i8++)
        {
            const int j4 = (int)(pi * g) + (int)i8;
            multiply_vector_named_accumulate(evaluator, total_sum, total_started, moved[i8], owner, (std::size_t)j4, 1,
                                             [&]() -> const vector<double> & { return P.select_one_vec[(std::size_t)j4]; });
        }
    }
    evaluator.rescale_to_next_inplace(total_sum);
    var = total_sum;

    if (!end)
    { // replicate into po copies
        if (zero_start)
            sum = ct_zero;
        vector<int> copy_steps;
        for (int u6 = 0; u6 < po; u6++)
            copy_steps.push_back((int)(-u6 * (n / po)));
        vector<Ciphertext> copies;
        rotated_copies(var, copy_steps, copies, evaluator, gal_keys);
        for (int u6 = 0; u6 < po; u6++)
        {
            if (u6 == 0 && !zero_start)
                sum = copies[(std::size_t)u6];
            else
                evaluator.add_inplace_reduced_error(sum, copies[(std::size_t)u6]);
        }
        var = sum;
    }
    cnn_out = TensorCipher(logn, ko, ho, wo, co, to, po, var);
}

void multiplexed_parallel_convolution_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, int co, int st, int fh, int fw,
                                           const vector<double> &data, vector<double> running_var,
                                           vector<double> constant_weight, double epsilon, CKKSEncoder &encoder,
                                           Encryptor &encryptor, Evaluator &evaluator, GaloisKeys &gal_keys,
                                           vector<Ciphertext> & /*cipher_pool: host-memory reuse, not needed here*/, bool end)
{
    // the reference rebuilds the weight and mask vectors on every call; so does this entry point (nothing is cached)
    ConvPlan plan = build_conv_plan(cnn_in, co, st, fh, fw, data, running_var, constant_weight, epsilon);
    multiplexed_parallel_convolution_planned(cnn_in, cnn_out, plan, encoder, encryptor, evaluator, gal_keys, end, false);
}

void multiplexed_parallel_batch_norm_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, vector<double> bias,
                                          vector<double> running_mean, vector<double> running_var, vector<double> weight,
                                          double epsilon, CKKSEncoder &encoder, Encryptor &encryptor, Evaluator &evaluator,
                                          double B, bool)
{
    multiplexed_parallel_batch_norm_named(cnn_in, cnn_out, bias, running_mean, running_var, weight, epsilon, encoder, encryptor,
                                          evaluator, B, nullptr, 0);
}

// The same with a NAME for the shift vector (owner, index): a layer's shift is a parameter of the network, so on the
// engine its encoding stays in HBM per (level, scale) instead of being rebuilt and encoded for every image.
void multiplexed_parallel_batch_norm_named(const TensorCipher &cnn_in, TensorCipher &cnn_out, const vector<double> &bias,
                                           const vector<double> &running_mean, const vector<double> &running_var,
                                           const vector<double> &weight, double epsilon, CKKSEncoder &encoder,
                                           Encryptor &encryptor, Evaluator &evaluator, double B, const void *owner,
                                           std::uint64_t index)
{
    const int ki = cnn_in.k(), hi = cnn_in.h(), wi = cnn_in.w(), ci = cnn_in.c(), ti = cnn_in.t(), pi = cnn_in.p(),
              logn = cnn_in.logn();
    if ((int)bias.size() != ci || (int)running_mean.size() != ci || (int)running_var.size() != ci || (int)weight.size() != ci)
        throw std::invalid_argument("the size of bias, running_mean, running_var, or weight are not correct");
    for (double v : running_var)
        if (v < 1e-16 && v > -1e-16)
            throw std::invalid_argument("the size of running_var is too small. nearly zero.");
    if ((long)hi * wi * ci > (1L << logn))
        throw std::invalid_argument("hi*wi*ci should not be larger than n");
    const long n = 1L << logn;
    if (n % pi != 0)
        throw std::out_of_range("n is not divisible by pi");

    // the convolution already multiplied by weight / sqrt(var + eps); what is left is the constant
    // (mean * weight / sqrt(var + eps) - bias) / B per channel, laid out like the tensor
    const long per_copy = n / pi, plane = (long)ki * ki * hi * wi, row = (long)ki * wi;
    vector<double> g;
    auto shift = [&]() -> const vector<double> & {
        g.assign((std::size_t)n, 0.0);
        for (long slot = 0; slot < n; slot++)
        {
            const long r = slot % per_copy;
            if (r >= plane * ti)
                continue;
            const long ch = (long)ki * ki * (r / plane) + ki * (((r % plane) / row) % ki) + (r % row) % ki;
            if (ch >= ci)
                continue;
            g[(std::size_t)slot] = (running_mean[(std::size_t)ch] * weight[(std::size_t)ch] /
                                        std::sqrt(running_var[(std::size_t)ch] + epsilon) -
                                    bias[(std::size_t)ch]) /
                                   B;
        }
        return g;
    };
    Plaintext plain;
    Ciphertext cipher_g, temp = cnn_in.cipher();
#ifdef B200CKKS_FACADE
    if (owner && !encrypt_constants())
    {
        evaluator.sub_vector_inplace_cached(temp, owner, index, 2, shift);
        cnn_out = TensorCipher(logn, ki, hi, wi, ci, ti, pi, temp);
        return;
    }
#else
    (void)owner;
    (void)index;
#endif
    shift();
    if (encrypt_constants())
    { // the reference's sequence (cnn_seal.cpp:566-570): a fresh top-level encryption of g, walked down by the subtraction
        encoder.encode(g, temp.scale(), plain);
        encryptor.encrypt(plain, cipher_g);
        evaluator.sub_inplace_reduced_error(temp, cipher_g);
    }
    else
    { // g as the plaintext it is, encoded at the ciphertext's own level and scale
        encoder.encode(g, temp.parms_id(), temp.scale(), plain);
        evaluator.sub_plain_inplace(temp, plain);
    }
    cnn_out = TensorCipher(logn, ki, hi, wi, ci, ti, pi, temp);
}

void ReLU_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, long comp_no, vector<int> deg, long alpha,
               vector<minicomp::Tree> &tree, double scaled_val, long scalingfactor, Encryptor &encryptor, Evaluator &evaluator,
               Decryptor &decryptor, CKKSEncoder &encoder, PublicKey &public_key, SecretKey &secret_key, RelinKeys &relin_keys,
               double)
{
    if ((long)cnn_in.h() * cnn_in.w() * cnn_in.c() > (1L << cnn_in.logn()))
        throw std::invalid_argument("hi*wi*ci should not be larger than n");
    Ciphertext temp = cnn_in.cipher();
    minimax_ReLU_seal(comp_no, deg, alpha, tree, scaled_val, scalingfactor, encryptor, evaluator, decryptor, encoder, public_key,
                      secret_key, relin_keys, temp, temp);
    cnn_out = TensorCipher(cnn_in.logn(), cnn_in.k(), cnn_in.h(), cnn_in.w(), cnn_in.c(), cnn_in.t(), cnn_in.p(), temp);
}

void cnn_add_seal(const TensorCipher &cnn1, const TensorCipher &cnn2, TensorCipher &destination, Evaluator &evaluator)
{
    if (cnn1.k() != cnn2.k() || cnn1.h() != cnn2.h() || cnn1.w() != cnn2.w() || cnn1.c() != cnn2.c() || cnn1.t() != cnn2.t() ||
        cnn1.p() != cnn2.p() || cnn1.logn() != cnn2.logn())
        throw std::invalid_argument("the parameters of cnn1 and cnn2 are not the same");
    Ciphertext a = cnn1.cipher(), b = cnn2.cipher();
    evaluator.add_inplace_reduced_error(a, b);
    destination = TensorCipher(cnn1.logn(), cnn1.k(), cnn1.h(), cnn1.w(), cnn1.c(), cnn1.t(), cnn1.p(), a);
}

void multiplexed_parallel_downsampling_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, Evaluator &evaluator,
                                            GaloisKeys &gal_keys)
{
    const int ki = cnn_in.k(), hi = cnn_in.h(), wi = cnn_in.w(), ci = cnn_in.c(), ti = cnn_in.t(), logn = cnn_in.logn();
    const long n = 1L << logn;
    const int ko = 2 * ki, ho = hi / 2, wo = wi / 2, to = ti / 2, co = 2 * ci;
    if (ti % 8 != 0)
        throw std::invalid_argument("ti is not multiple of 8");
    if (hi % 2 != 0)
        throw std::invalid_argument("hi is not even");
    if (wi % 2 != 0)
        throw std::invalid_argument("wi is not even");
    const int po = copies_that_fit(n, (long)ko * ko * ho * wo * to);
    if (n % po != 0)
        throw std::out_of_range("n is not divisible by po");

    Ciphertext ct = cnn_in.cipher(), sum, temp;
    const long plane = (long)ki * ki * hi * wi, row = (long)ki * wi;
    Accumulator acc{ evaluator, sum };
    vector<double> mask((std::size_t)n);
    for (int w1 = 0; w1 < ki; w1++)
        for (int w2 = 0; w2 < ti; w2++)
        {
            // keep, inside channel group w2, sub-row w1 of the cells at even (y, x)
            for (long slot = 0; slot < n; slot++)
            {
                const long yy = (slot % plane) / row, xx = slot % row, group = slot / plane;
                mask[(std::size_t)slot] = (slot < plane * ti && (yy / ki) % 2 == 0 && (xx / ki) % 2 == 0 && yy % ki == w1 && group == w2) ? 1.0 : 0.0;
            }
            temp = ct;
            evaluator.multiply_vector_inplace_reduced_error(temp, mask);
            const int lin = ki * w2 + w1;
            const int w3 = (lin % (2 * ko)) / 2, w4 = lin % 2, w5 = lin / (2 * ko);
            memory_save_rotate(temp, temp,
                               (int)(plane * w2 + row * w1 - (long)ko * ko * ho * wo * w5 - (long)ko * wo * w3 - ki * w4 -
                                     (long)ko * ko * ho * wo * (ti / 8)),
                               evaluator, gal_keys);
            acc.add(temp);
        }
    evaluator.rescale_to_next_inplace(sum);
    ct = sum;

    for (int u6 = 1; u6 < po; u6++)
    {
        rotated_copy(ct, temp, (int)(-(n / po) * u6), evaluator, gal_keys);
        evaluator.add_inplace_reduced_error(sum, temp);
    }
    cnn_out = TensorCipher(logn, ko, ho, wo, co, to, po, sum);
}

void averagepooling_seal_scale(const TensorCipher &cnn_in, TensorCipher &cnn_out, Evaluator &evaluator, GaloisKeys &gal_keys,
                               double B)
{
    const int ki = cnn_in.k(), hi = cnn_in.h(), wi = cnn_in.w(), ci = cnn_in.c(), ti = cnn_in.t(), logn = cnn_in.logn();
    if (log2_long(hi) == -1)
        throw std::invalid_argument("hi is not power of two");
    if (log2_long(wi) == -1)
        throw std::invalid_argument("wi is not power of two");
    const long n = 1L << logn;
    Ciphertext ct = cnn_in.cipher(), temp, sum;

    // sum over x, then over y
    for (int x = 0; x < log2_long(wi); x++)
    {
        rotated_copy(ct, temp, (int)pow2(x) * ki, evaluator, gal_keys);
        evaluator.add_inplace_reduced_error(ct, temp);
    }
    for (int x = 0; x < log2_long(hi); x++)
    {
        rotated_copy(ct, temp, (int)pow2(x) * ki * ki * wi, evaluator, gal_keys);
        evaluator.add_inplace_reduced_error(ct, temp);
    }
    // gather the per-channel sums into consecutive slots, scaled by B / (h w) (undoes the input's 1/B)
    Accumulator acc{ evaluator, sum };
    vector<double> select_one((std::size_t)n);
    for (int s = 0; s < ki; s++)
        for (int u = 0; u < ti; u++)
        {
            const int p = ki * u + s;
            rotated_copy(ct, temp, -p * ki + ki * ki * hi * wi * u + ki * wi * s, evaluator, gal_keys);
            std::fill(select_one.begin(), select_one.end(), 0.0);
            for (int i = 0; i < ki; i++)
                select_one[(std::size_t)(p * ki + i)] = B / static_cast<double>(hi * wi);
            evaluator.multiply_vector_inplace_reduced_error(temp, select_one);
            acc.add(temp);
        }
    evaluator.rescale_to_next_inplace(sum);
    cnn_out = TensorCipher(logn, 1, 1, 1, ci, ti, 1, sum);
}

void matrix_multiplication_seal(const TensorCipher &cnn_in, TensorCipher &cnn_out, vector<double> matrix, vector<double> bias,
                                int q, int r, Evaluator &evaluator, GaloisKeys &gal_keys, const void *owner)
{
    const int logn = cnn_in.logn();
    if ((int)matrix.size() != q * r)
        throw std::invalid_argument("the size of matrix is not q*r");
    if ((int)bias.size() != q)
        throw std::invalid_argument("the size of bias is not q");
    const long n = 1L << logn;
    // diagonal s of the q x r matrix: W[s][i] = matrix[i][j] with i - j + r - 1 = s.
    // (The reference builds a bias vector here as well but never adds it - cnn_seal.cpp:759,763 - so neither do we.)
    vector<vector<double>> W((std::size_t)(q + r - 1), vector<double>((std::size_t)n, 0.0));
    for (int i = 0; i < q; i++)
        for (int j = 0; j < r; j++)
            W[(std::size_t)(i - j + r - 1)][(std::size_t)i] = matrix[(std::size_t)(i * r + j)];

    Ciphertext ct = cnn_in.cipher(), temp, sum;
#ifdef B200CKKS_FACADE
    if (owner && !encrypt_constants())
    {
        // all q + r - 1 rotations are rotations of ONE ciphertext (shared decomposition), the diagonals are parameters
        // of the network (encoded once, kept in HBM), and the sum of products is one pass
        vector<int> steps;
        for (int s = 0; s < q + r - 1; s++)
            steps.push_back(r - 1 - s);
        vector<Ciphertext> rot;
        rotated_copies(ct, steps, rot, evaluator, gal_keys);
        vector<const Ciphertext *> terms;
        vector<std::uint64_t> ids;
        for (int s = 0; s < q + r - 1; s++)
        {
            terms.push_back(&rot[(std::size_t)s]);
            ids.push_back((std::uint64_t)s);
        }
        multiply_vector_named_sum(evaluator, sum, terms, owner, ids, 3,
                                  [&](std::uint64_t idx) -> const vector<double> & { return W[(std::size_t)idx]; });
        evaluator.rescale_to_next_inplace(sum);
        cnn_out = TensorCipher(logn, cnn_in.k(), cnn_in.h(), cnn_in.w(), cnn_in.c(), cnn_in.t(), cnn_in.p(), sum);
        return;
    }
#else
    (void)owner;
#endif
    Accumulator acc{ evaluator, sum };
    for (int s = 0; s < q + r - 1; s++)
    {
        rotated_copy(ct, temp, r - 1 - s, evaluator, gal_keys);
        evaluator.multiply_vector_inplace_reduced_error(temp, W[(std::size_t)s]);
        acc.add(temp);
    }
    evaluator.rescale_to_next_inplace(sum);
    cnn_out = TensorCipher(logn, cnn_in.k(), cnn_in.h(), cnn_in.w(), cnn_in.c(), cnn_in.t(), cnn_in.p(), sum);
}
