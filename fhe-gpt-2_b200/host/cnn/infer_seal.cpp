// cnn/infer_seal.cpp - see infer_seal.h.
#include "cnn/infer_seal.h"
#include "common/cached.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <cmath>
#include <fstream>
#include <iostream>
#include <stdexcept>

using namespace seal;
using std::string;
using std::vector;

int resnet_end_num(std::size_t layer_num)
{
    switch (layer_num)
    {
    case 20: return 2;
    case 32: return 4;
    case 44: return 6;
    case 56: return 8;
    case 110: return 17;
    default: throw std::invalid_argument("layer_num is not correct");
    }
}

void resnet_conv_shape(std::size_t layer_num, std::size_t index, int &ci, int &co)
{
    const int blocks = resnet_end_num(layer_num) + 1;
    if (index == 0)
    {
        ci = 3;
        co = 16;
        return;
    }
    const int pos = (int)index - 1;            // 0 .. 6 * blocks - 1
    const int stage = pos / (2 * blocks);      // 0, 1, 2 -> 16, 32, 64 channels
    const bool first_of_stage = pos % (2 * blocks) == 0;
    co = 16 << stage;
    ci = (first_of_stage && stage > 0) ? co / 2 : co;
}

namespace
{
    void read_values(const string &path, std::size_t count, vector<double> &out)
    {
        std::ifstream in(path);
        if (!in.is_open())
            throw std::runtime_error("file is not open: " + path);
        out.clear();
        double v;
        for (std::size_t i = 0; i < count; i++)
        {
            if (!(in >> v))
                throw std::runtime_error("file is too short: " + path);
            out.push_back(v);
        }
    }
} // namespace

void import_parameters_cifar10(vector<double> &linear_weight, vector<double> &linear_bias, vector<vector<double>> &conv_weight,
                               vector<vector<double>> &bn_bias, vector<vector<double>> &bn_running_mean,
                               vector<vector<double>> &bn_running_var, vector<vector<double>> &bn_weight, std::size_t layer_num,
                               std::size_t end_num, const string &dir)
{
    if ((int)end_num != resnet_end_num(layer_num))
        throw std::invalid_argument("layer number is not valid");
    const string base = dir + "/resnet" + std::to_string(layer_num) + "_new/";
    const std::size_t layers = layer_num - 1;
    conv_weight.assign(layers, {});
    bn_bias.assign(layers, {});
    bn_running_mean.assign(layers, {});
    bn_running_var.assign(layers, {});
    bn_weight.assign(layers, {});
    for (std::size_t idx = 0; idx < layers; idx++)
    {
        int ci, co;
        resnet_conv_shape(layer_num, idx, ci, co);
        string conv, bn;
        if (idx == 0)
        {
            conv = "conv1";
            bn = "bn1";
        }
        else
        {
            const std::size_t pos = idx - 1, blocks = end_num + 1;
            const string block = "layer" + std::to_string(pos / (2 * blocks) + 1) + "_" + std::to_string((pos % (2 * blocks)) / 2);
            conv = block + "_conv" + std::to_string(pos % 2 + 1);
            bn = block + "_bn" + std::to_string(pos % 2 + 1);
        }
        read_values(base + conv + "_weight.txt", (std::size_t)(9 * ci * co), conv_weight[idx]);
        read_values(base + bn + "_bias.txt", (std::size_t)co, bn_bias[idx]);
        read_values(base + bn + "_running_mean.txt", (std::size_t)co, bn_running_mean[idx]);
        read_values(base + bn + "_running_var.txt", (std::size_t)co, bn_running_var[idx]);
        read_values(base + bn + "_weight.txt", (std::size_t)co, bn_weight[idx]);
    }
    read_values(base + "linear_weight.txt", 10 * 64, linear_weight);
    read_values(base + "linear_bias.txt", 10, linear_bias);
}

void import_parameters_cifar100(ResNetParameters &p, std::size_t layer_num, const string &dir)
{
    // infer_seal.cpp:108-250: same file names as CIFAR-10 under <dir>/resnet32_cifar100, plus the two shortcuts
    const std::size_t end_num = (std::size_t)resnet_end_num(layer_num), layers = layer_num - 1;
    const string base = dir + "/resnet" + std::to_string(layer_num) + "_cifar100/";
    p.conv_weight.assign(layers, {});
    p.bn_bias.assign(layers, {});
    p.bn_running_mean.assign(layers, {});
    p.bn_running_var.assign(layers, {});
    p.bn_weight.assign(layers, {});
    for (std::size_t idx = 0; idx < layers; idx++)
    {
        int ci, co;
        resnet_conv_shape(layer_num, idx, ci, co);
        string conv = "conv1", bn = "bn1";
        if (idx > 0)
        {
            const std::size_t pos = idx - 1, blocks = end_num + 1;
            const string block = "layer" + std::to_string(pos / (2 * blocks) + 1) + "_" + std::to_string((pos % (2 * blocks)) / 2);
            conv = block + "_conv" + std::to_string(pos % 2 + 1);
            bn = block + "_bn" + std::to_string(pos % 2 + 1);
        }
        read_values(base + conv + "_weight.txt", (std::size_t)(9 * ci * co), p.conv_weight[idx]);
        read_values(base + bn + "_bias.txt", (std::size_t)co, p.bn_bias[idx]);
        read_values(base + bn + "_running_mean.txt", (std::size_t)co, p.bn_running_mean[idx]);
        read_values(base + bn + "_running_var.txt", (std::size_t)co, p.bn_running_var[idx]);
        read_values(base + bn + "_weight.txt", (std::size_t)co, p.bn_weight[idx]);
    }
    p.shortcut_weight.assign(2, {});
    p.shortcut_bn_bias.assign(2, {});
    p.shortcut_bn_mean.assign(2, {});
    p.shortcut_bn_var.assign(2, {});
    p.shortcut_bn_weight.assign(2, {});
    for (std::size_t j = 0; j < 2; j++)
    {
        const std::size_t ci = 16u << j, co = 32u << j;
        const string sc = base + "layer" + std::to_string(j + 2) + "_0_shortcut_";
        read_values(sc + "0_weight.txt", ci * co, p.shortcut_weight[j]);
        read_values(sc + "1_bias.txt", co, p.shortcut_bn_bias[j]);
        read_values(sc + "1_running_mean.txt", co, p.shortcut_bn_mean[j]);
        read_values(sc + "1_running_var.txt", co, p.shortcut_bn_var[j]);
        read_values(sc + "1_weight.txt", co, p.shortcut_bn_weight[j]);
    }
    read_values(base + "linear_weight.txt", 100 * 64, p.linear_weight);
    read_values(base + "linear_bias.txt", 100, p.linear_bias);
}

const vector<int> &resnet_rotation_kinds()
{
    // the steps memory_save_rotate is asked for by the convolutions, down-samplings, pooling and the FC layer of the
    // three stages (infer_seal.cpp:345-362): data of the packing scheme, identical for every depth
    static const vector<int> kinds = {
        0,     1,     2,     3,     4,     5,     6,     7,     8,     9,     10,    11,    12,    13,    14,    15,    16,
        17,    18,    19,    20,    21,    22,    23,    24,    25,    26,    27,    28,    29,    30,    31,    32,    33,
        56,    62,    63,    64,    66,    84,    124,   128,   132,   256,   512,   959,   960,   990,   991,   1008,  1023,
        1024,  1036,  1064,  1092,  1952,  1982,  1983,  2016,  2044,  2047,  2048,  2072,  2078,  2100,  3007,  3024,  3040,
        3052,  3070,  3071,  3072,  3080,  3108,  4031,  4032,  4062,  4063,  4095,  4096,  5023,  5024,  5054,  5055,  5087,
        5118,  5119,  5120,  6047,  6078,  6079,  6111,  6112,  6142,  6143,  6144,  7071,  7102,  7103,  7135,  7166,  7167,
        7168,  8095,  8126,  8127,  8159,  8190,  8191,  8192,  9149,  9183,  9184,  9213,  9215,  9216,  10173, 10207, 10208,
        10237, 10239, 10240, 11197, 11231, 11232, 11261, 11263, 11264, 12221, 12255, 12256, 12285, 12287, 12288, 13214, 13216,
        13246, 13278, 13279, 13280, 13310, 13311, 13312, 14238, 14240, 14270, 14302, 14303, 14304, 14334, 14335, 15262, 15264,
        15294, 15326, 15327, 15328, 15358, 15359, 15360, 16286, 16288, 16318, 16350, 16351, 16352, 16382, 16383, 16384, 17311,
        17375, 18335, 18399, 18432, 19359, 19423, 20383, 20447, 20480, 21405, 21406, 21437, 21469, 21470, 21471, 21501, 21504,
        22429, 22430, 22461, 22493, 22494, 22495, 22525, 22528, 23453, 23454, 23485, 23517, 23518, 23519, 23549, 24477, 24478,
        24509, 24541, 24542, 24543, 24573, 24576, 25501, 25565, 25568, 25600, 26525, 26589, 26592, 26624, 27549, 27613, 27616,
        27648, 28573, 28637, 28640, 28672, 29600, 29632, 29664, 29696, 30624, 30656, 30688, 30720, 31648, 31680, 31712, 31743,
        31744, 31774, 32636, 32640, 32644, 32672, 32702, 32704, 32706, 32735, 32736, 32737, 32759, 32760, 32761, 32762, 32763,
        32764, 32765, 32766, 32767
    };
    return kinds;
}

vector<int> ResNetCifar10::coeff_bit_vec()
{
    vector<int> bits;
    bits.push_back(logq);
    for (int i = 0; i < remaining_level; i++)
        bits.push_back(logp);
    for (int i = 0; i < boot_level; i++)
        bits.push_back(logq);
    bits.push_back(log_special_prime);
    return bits;
}

ResNetCifar10::ResNetCifar10(std::size_t layer_num, ResNetParameters parameters, SEALContext &context, KeyGenerator &keygen,
                             CKKSEncoder &encoder, Encryptor &encryptor, Decryptor &decryptor, Evaluator &evaluator,
                             PublicKey &public_key, SecretKey &secret_key, RelinKeys &relin_keys, GaloisKeys &gal_keys,
                             ResNetVariant variant)
    : B(variant.B), layer_num_(layer_num), end_num_(resnet_end_num(layer_num)), variant_(variant), w_(std::move(parameters)),
      context_(context), keygen_(keygen),
      encoder_(encoder), encryptor_(encryptor), decryptor_(decryptor), evaluator_(evaluator), public_key_(public_key),
      secret_key_(secret_key), relin_keys_(relin_keys), gal_keys_(gal_keys)
{
    const std::size_t layers = layer_num - 1;
    if (w_.conv_weight.size() != layers || w_.bn_bias.size() != layers || w_.bn_running_mean.size() != layers ||
        w_.bn_running_var.size() != layers || w_.bn_weight.size() != layers ||
        w_.linear_weight.size() != (std::size_t)variant_.classes * 64 || w_.linear_bias.size() != (std::size_t)variant_.classes)
        throw std::invalid_argument("parameter lists do not match the network depth");
    if (variant_.shortcut_conv &&
        (w_.shortcut_weight.size() != 2 || w_.shortcut_bn_bias.size() != 2 || w_.shortcut_bn_mean.size() != 2 ||
         w_.shortcut_bn_var.size() != 2 || w_.shortcut_bn_weight.size() != 2))
        throw std::invalid_argument("two shortcut convolutions are required");
    conv_plans_.resize(layers + 2); // the last two slots: shortcut convolutions of stages 2 and 3
    for (long i = 0; i < comp_no; i++)
    {
        minicomp::Tree tr;
        upgrade_oddbaby(deg_[(std::size_t)i], tr);
        tree_.push_back(tr);
    }
    const int total_level = remaining_level + boot_level;
    const double scale = std::pow(2.0, logp);
    const long sparse_logn[3] = { 14, 13, 12 }; // 16 / 32 / 64-channel stages
    for (int i = 0; i < 3; i++)
    {
        boot_[i] = new Bootstrapper(loge, sparse_logn[i], logN - 1, total_level, scale, boundary_K, boot_deg, scale_factor,
                                    inverse_deg, context_, keygen_, encoder_, encryptor_, decryptor_, evaluator_, relin_keys_,
                                    gal_keys_);
        boot_[i]->prepare_mod_polynomial();
    }
}

ResNetCifar10::~ResNetCifar10()
{
    for (auto &p : conv_plans_)
        if (p)
            forget_named(evaluator_, p.get());
    forget_named(evaluator_, this); // the batch-norm shifts cached under this object's name
    forget_named(evaluator_, &w_.linear_weight);
    for (auto *b : boot_)
        delete b;
}

vector<int> ResNetCifar10::galois_steps() const
{
    vector<int> steps;
    steps.push_back(0);
    for (int i = 0; i < logN - 1; i++)
        steps.push_back(1 << i);
    for (int rot : resnet_rotation_kinds())
        if (std::find(steps.begin(), steps.end(), rot) == steps.end())
            steps.push_back(rot);
    // the fully connected layer rotates by 63 .. -(classes - 1); the reference's list stops at -9 (ten classes)
    for (int k = 10; k < variant_.classes; k++)
        steps.push_back((1 << logn) - k);
    for (auto *b : boot_)
        b->addLeftRotKeys_Linear_to_vector_3(steps);
    return steps;
}

void ResNetCifar10::prepare()
{
    for (auto *b : boot_)
    {
        b->slot_vec.push_back(b->logn);
        b->generate_LT_coefficient_3();
    }
    prepared_ = true;
}

vector<double> ResNetCifar10::infer(const vector<double> &image_in, vector<ResNetTraceRow> *trace)
{
    return decrypt_logits(infer_encrypted(encrypt_image(image_in), trace));
}

TensorCipher ResNetCifar10::encrypt_image(const vector<double> &image_in)
{
    if (image_in.size() != 32 * 32 * 3)
        throw std::invalid_argument("image must hold 32*32*3 values");
    const long n = 1L << logn, init_p = 8;
    // pack: 8 copies of the 3 x 32 x 32 image, values divided by B so that activations stay in [-1, 1]
    vector<double> image((std::size_t)n, 0.0);
    for (std::size_t i = 0; i < image_in.size(); i++)
        image[i] = image_in[i];
    for (long i = n / init_p; i < n; i++)
        image[(std::size_t)i] = image[(std::size_t)(i % (n / init_p))];
    for (auto &v : image)
        v /= B;
    return TensorCipher(logn, 1, 32, 32, 3, 3, init_p, image, encryptor_, encoder_, logq);
}

vector<double> ResNetCifar10::decrypt_logits(const TensorCipher &output)
{
    Plaintext plain;
    decryptor_.decrypt(output.cipher_ref(), plain);
    vector<std::complex<double>> slots;
    encoder_.decode(plain, slots);
    vector<double> logits((std::size_t)variant_.classes);
    for (std::size_t i = 0; i < logits.size(); i++)
        logits[i] = slots[i].real();
    return logits;
}

TensorCipher ResNetCifar10::infer_encrypted(const TensorCipher &input, vector<ResNetTraceRow> *trace)
{
    if (!prepared_)
        throw std::logic_error("ResNetCifar10::prepare() must be called after the Galois keys were created");
    const double epsilon = 0.00001;

    auto t_prev = std::chrono::high_resolution_clock::now();
    auto log_op = [&](int op, const TensorCipher &t) {
        if (!trace)
            return;
#ifdef B200CKKS_FACADE
        context_.sync();
#endif
        auto now = std::chrono::high_resolution_clock::now();
        const Ciphertext &c = t.cipher_ref();
        trace->push_back({ op, (int)context_.get_context_data(c.parms_id())->chain_index(), c.scale(),
                           std::chrono::duration<double, std::milli>(now - t_prev).count() });
        t_prev = now;
    };

    TensorCipher cnn = input, temp;
    {
        Ciphertext ctxt = cnn.cipher();
        for (int i = 0; i < boot_level - 3; i++)
            evaluator_.mod_switch_to_next_inplace(ctxt);
        cnn.set_ciphertext(ctxt);
    }
    t_prev = std::chrono::high_resolution_clock::now();

    // weights and masks of a layer are the same for every image: built once, their encodings stay in HBM
    auto planned_conv = [&](TensorCipher &t, std::size_t slot, int co, int st, int f, const vector<double> &weight,
                            const vector<double> &var, const vector<double> &gamma) {
        std::unique_ptr<ConvPlan> &plan = conv_plans_[slot];
        {
            std::lock_guard<std::mutex> guard(plan_mu_); // images may run on several host threads (infer_seal.cpp:404)
            if (!plan)
                plan = std::make_unique<ConvPlan>(build_conv_plan(t, co, st, f, f, weight, var, gamma, epsilon));
        }
        multiplexed_parallel_convolution_planned(t, t, *plan, encoder_, encryptor_, evaluator_, gal_keys_);
        log_op(0, t);
    };
    auto conv = [&](int stage, int co, int st) {
        planned_conv(cnn, (std::size_t)stage, co, st, 3, w_.conv_weight[(std::size_t)stage], w_.bn_running_var[(std::size_t)stage],
                     w_.bn_weight[(std::size_t)stage]);
    };
    auto bn = [&](int stage) {
        multiplexed_parallel_batch_norm_named(cnn, cnn, w_.bn_bias[(std::size_t)stage], w_.bn_running_mean[(std::size_t)stage],
                                              w_.bn_running_var[(std::size_t)stage], w_.bn_weight[(std::size_t)stage], epsilon,
                                              encoder_, encryptor_, evaluator_, B, this, (std::uint64_t)stage);
        log_op(1, cnn);
    };
    auto relu = [&]() {
        ReLU_seal(cnn, cnn, comp_no, deg_, alpha, tree_, scaled_val, logp, encryptor_, evaluator_, decryptor_, encoder_,
                  public_key_, secret_key_, relin_keys_, B);
        log_op(2, cnn);
    };
    auto bootstrap = [&](int j) {
        Ciphertext ctxt = cnn.cipher(), rtn;
        boot_[j]->bootstrap_real_3(rtn, ctxt);
        cnn.set_ciphertext(rtn);
        log_op(3, cnn);
    };

    // layer 0; its convolution runs at a 2^51 scale, bring the result back to exactly 2^46
    conv(0, 16, 1);
    {
        const auto &modulus = util::iter(context_.first_context_data()->parms().coeff_modulus());
        Ciphertext ctxt = cnn.cipher();
        const std::size_t cur_level = ctxt.coeff_modulus_size();
        Plaintext scaler;
        const double scale_change = std::pow(2.0, 46) * ((double)modulus[cur_level - 1].value()) / ctxt.scale();
        encoder_.encode(1.0, scale_change, scaler);
        evaluator_.mod_switch_to_inplace(scaler, ctxt.parms_id());
        evaluator_.multiply_plain_inplace(ctxt, scaler);
        evaluator_.rescale_to_next_inplace(ctxt);
        ctxt.scale() = std::pow(2.0, 46);
        cnn.set_ciphertext(ctxt);
    }
    bn(0);
    relu();

    for (int j = 0; j < 3; j++)
    {
        const int co = 16 << j;
        for (int k = 0; k <= end_num_; k++)
        {
            int stage = 2 * ((end_num_ + 1) * j + k) + 1;
            temp = cnn;
            conv(stage, co, (j >= 1 && k == 0) ? 2 : 1);
            bn(stage);
            bootstrap(j);
            relu();

            stage++;
            conv(stage, co, 1);
            bn(stage);
            if (j >= 1 && k == 0 && variant_.shortcut_conv)
            {
                // CIFAR-100: 1x1 stride-2 convolution + batch norm on the shortcut (infer_seal.cpp:826-831)
                const std::size_t sc = (std::size_t)j - 1;
                planned_conv(temp, layer_num_ - 1 + sc, co, 2, 1, w_.shortcut_weight[sc], w_.shortcut_bn_var[sc],
                             w_.shortcut_bn_weight[sc]);
                multiplexed_parallel_batch_norm_named(temp, temp, w_.shortcut_bn_bias[sc], w_.shortcut_bn_mean[sc],
                                                      w_.shortcut_bn_var[sc], w_.shortcut_bn_weight[sc], epsilon, encoder_,
                                                      encryptor_, evaluator_, B, this, (std::uint64_t)(1000 + sc));
                log_op(1, temp);
            }
            else if (j >= 1 && k == 0)
            {
                multiplexed_parallel_downsampling_seal(temp, temp, evaluator_, gal_keys_);
                log_op(5, temp);
            }
            cnn_add_seal(temp, cnn, cnn, evaluator_);
            log_op(4, cnn);
            bootstrap(j);
            relu();
        }
    }
    averagepooling_seal_scale(cnn, cnn, evaluator_, gal_keys_, B);
    log_op(6, cnn);
    matrix_multiplication_seal(cnn, cnn, w_.linear_weight, w_.linear_bias, variant_.classes, 64, evaluator_, gal_keys_,
                               &w_.linear_weight);
    log_op(7, cnn);
    return cnn;
}

// ------------------------------------------------------------------------------------------ the reference's driver
namespace
{
    bool readable(const string &path)
    {
        std::ifstream f(path);
        return f.is_open();
    }

    // random-init weights of the architecture (He-style convolutions, near-identity batch-norm statistics)
    ResNetParameters random_parameters(std::size_t layer_num, unsigned long long seed, const ResNetVariant &variant = {})
    {
        ResNetParameters p;
        unsigned long long state = seed * 0x9E3779B97F4A7C15ull + 0x5EA1ull;
        auto uni = [&state]() { // xorshift64*, uniform in (0,1)
            state ^= state >> 12;
            state ^= state << 25;
            state ^= state >> 27;
            return ((state * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0) + 1e-17;
        };
        auto normal = [&uni]() { return std::sqrt(-2.0 * std::log(uni())) * std::cos(2.0 * M_PI * uni()); };
        const std::size_t layers = layer_num - 1;
        for (std::size_t i = 0; i < layers; i++)
        {
            int ci, co;
            resnet_conv_shape(layer_num, i, ci, co);
            vector<double> w((std::size_t)(9 * ci * co)), b((std::size_t)co), m((std::size_t)co), v((std::size_t)co), g((std::size_t)co);
            const double sd = 0.5 * std::sqrt(2.0 / (9.0 * ci));
            for (auto &x : w)
                x = sd * normal();
            for (int c = 0; c < co; c++)
            {
                b[(std::size_t)c] = 0.1 * normal();
                m[(std::size_t)c] = 0.1 * normal();
                v[(std::size_t)c] = 0.5 + uni();
                g[(std::size_t)c] = 0.5 + 0.5 * uni();
            }
            p.conv_weight.push_back(w);
            p.bn_bias.push_back(b);
            p.bn_running_mean.push_back(m);
            p.bn_running_var.push_back(v);
            p.bn_weight.push_back(g);
        }
        p.linear_weight.resize((std::size_t)variant.classes * 64);
        for (auto &x : p.linear_weight)
            x = 0.3 * normal();
        p.linear_bias.assign((std::size_t)variant.classes, 0.0);
        if (variant.shortcut_conv)
            for (int j = 0; j < 2; j++)
            {
                const int ci = 16 << j, co = 32 << j;
                vector<double> w((std::size_t)(ci * co)), b((std::size_t)co), m((std::size_t)co), v((std::size_t)co), g((std::size_t)co);
                for (auto &x : w)
                    x = 0.5 * std::sqrt(2.0 / ci) * normal();
                for (int c = 0; c < co; c++)
                {
                    b[(std::size_t)c] = 0.1 * normal();
                    m[(std::size_t)c] = 0.1 * normal();
                    v[(std::size_t)c] = 0.5 + uni();
                    g[(std::size_t)c] = 0.5 + 0.5 * uni();
                }
                p.shortcut_weight.push_back(w);
                p.shortcut_bn_bias.push_back(b);
                p.shortcut_bn_mean.push_back(m);
                p.shortcut_bn_var.push_back(v);
                p.shortcut_bn_weight.push_back(g);
            }
        return p;
    }

    const char *op_title(int op)
    {
        static const char *t[] = { "multiplexed parallel convolution...", "multiplexed parallel batch normalization...",
                                   "approximate ReLU...", "bootstrapping...", "cipher add...",
                                   "multiplexed parallel downsampling...", "average pooling...", "fully connected layer..." };
        return t[op];
    }
} // namespace

static void resnet_driver(std::size_t layer_num, std::size_t start_image_id, std::size_t end_image_id, const string &result_dir,
                          const string &weights_dir, const string &images_dir, const ResNetVariant &variant);

void ResNet_cifar10_seal_sparse(std::size_t layer_num, std::size_t start_image_id, std::size_t end_image_id,
                                const string &result_dir, const string &weights_dir, const string &images_dir)
{
    resnet_driver(layer_num, start_image_id, end_image_id, result_dir, weights_dir, images_dir, ResNetVariant::cifar10());
}

void ResNet_cifar100_seal_sparse(std::size_t layer_num, std::size_t start_image_id, std::size_t end_image_id,
                                 const string &result_dir, const string &weights_dir, const string &images_dir)
{
    if (layer_num != 32)
        throw std::invalid_argument("layer number is not correct"); // infer_seal.cpp:615
    resnet_driver(layer_num, start_image_id, end_image_id, result_dir, weights_dir, images_dir, ResNetVariant::cifar100());
}

static void resnet_driver(std::size_t layer_num, std::size_t start_image_id, std::size_t end_image_id, const string &result_dir,
                          const string &weights_dir, const string &images_dir, const ResNetVariant &variant)
{
    const int end_num = resnet_end_num(layer_num);
    const string dataset = "cifar" + std::to_string(variant.classes);

    std::cout << "Setting Parameters" << std::endl;
    EncryptionParameters parms(scheme_type::ckks);
    const std::size_t poly_modulus_degree = std::size_t(1) << ResNetCifar10::logN;
    parms.set_poly_modulus_degree(poly_modulus_degree);
    parms.set_coeff_modulus(CoeffModulus::Create(poly_modulus_degree, ResNetCifar10::coeff_bit_vec()));
    parms.set_secret_key_hamming_weight(192);

    SEALContext context(parms);
    KeyGenerator keygen(context);
    PublicKey public_key;
    keygen.create_public_key(public_key);
    SecretKey secret_key = keygen.secret_key();
    RelinKeys relin_keys;
    keygen.create_relin_keys(relin_keys);
    GaloisKeys gal_keys;
    CKKSEncoder encoder(context);
    Encryptor encryptor(context, public_key);
    Evaluator evaluator(context, encoder);
    Decryptor decryptor(context, secret_key);

    ResNetParameters parameters;
    const string probe = weights_dir + "/resnet" + std::to_string(layer_num) + (variant.shortcut_conv ? "_cifar100" : "_new") +
                         "/conv1_weight.txt";
    const bool pretrained = readable(probe);
    if (pretrained && variant.shortcut_conv)
        import_parameters_cifar100(parameters, layer_num, weights_dir);
    else if (pretrained)
        import_parameters_cifar10(parameters.linear_weight, parameters.linear_bias, parameters.conv_weight, parameters.bn_bias,
                                  parameters.bn_running_mean, parameters.bn_running_var, parameters.bn_weight, layer_num,
                                  (std::size_t)end_num, weights_dir);
    else
    {
        std::cout << "no pretrained parameters at " << probe << ": random-init weights" << std::endl;
        parameters = random_parameters(layer_num, 0, variant);
    }

    std::cout << "Generating Optimal Minimax Polynomials..." << std::endl;
    ResNetCifar10 net(layer_num, parameters, context, keygen, encoder, encryptor, decryptor, evaluator, public_key, secret_key,
                      relin_keys, gal_keys, variant);
    std::cout << "Adding Bootstrapping Keys..." << std::endl;
    keygen.create_galois_keys(net.galois_steps(), gal_keys);
    std::cout << "Generating Linear Transformation Coefficients..." << std::endl;
    net.prepare();

    const string stem = result_dir + "/resnet" + std::to_string(layer_num) + "_" + dataset + "_";
    std::ofstream out_share(stem + "label_" + std::to_string(start_image_id) + "_" + std::to_string(end_image_id));
    const bool have_images = readable(images_dir + "/test_values.txt");
    if (!have_images)
        std::cout << "no " << images_dir << "/test_values.txt: synthetic images, labels not checked" << std::endl;

    // the reference's `#pragma omp parallel for num_threads(50)` over images (infer_seal.cpp:404): here
    // $B200CKKS_IMAGES_IN_FLIGHT host threads (default 1), each with its own CUDA stream, over the shared keys
    int in_flight = 1;
    if (const char *e = std::getenv("B200CKKS_IMAGES_IN_FLIGHT"))
        in_flight = std::max(1, std::atoi(e));
    in_flight = (int)std::min<std::size_t>((std::size_t)in_flight, end_image_id - start_image_id + 1);
    if (in_flight > 1)
    { // the first image alone: it builds the convolution plans, generates the level-pruned keys and fills the cache
        vector<double> warm(32 * 32 * 3, 0.0);
        net.infer(warm, nullptr);
    }
#ifdef B200CKKS_FACADE
    context.sync(); // set-up ran on this thread's stream; the image threads start from a drained device
#endif
    std::mutex report_mu;
    std::atomic<std::size_t> next_image{ start_image_id };
    std::exception_ptr failure;
    auto all_start = std::chrono::high_resolution_clock::now();
    auto image_loop = [&]() {
    try
    {
    for (std::size_t image_id = next_image++; image_id <= end_image_id; image_id = next_image++)
    {
        vector<double> image(32 * 32 * 3);
        int image_label = -1;
        if (have_images)
        {
            std::ifstream in(images_dir + "/test_values.txt");
            double val;
            for (std::size_t i = 0; i < 32 * 32 * 3 * image_id; i++)
                in >> val;
            for (auto &v : image)
                in >> v;
            std::ifstream in_label(images_dir + "/test_label.txt");
            for (std::size_t i = 0; i <= image_id; i++)
                in_label >> image_label;
        }
        else
        {
            unsigned long long state = (image_id + 1) * 0xD1B54A32D192ED03ull;
            auto uni = [&state]() {
                state ^= state >> 12;
                state ^= state << 25;
                state ^= state >> 27;
                return ((state * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0) + 1e-17;
            };
            for (auto &v : image)
                v = std::max(-2.5, std::min(2.5, std::sqrt(-2.0 * std::log(uni())) * std::cos(2.0 * M_PI * uni())));
        }

        vector<ResNetTraceRow> trace;
        auto t0 = std::chrono::high_resolution_clock::now();
        vector<double> logits = net.infer(image, &trace);
        const double total_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();

        std::ofstream output(stem + "image" + std::to_string(image_id) + ".txt");
        // the reference opens a "layer k" section at every convolution (infer_seal.cpp:465-536)
        int layer = 0;
        for (const auto &r : trace)
        {
            if (r.op == 0 || r.op == 6)
                output << "layer " << (r.op == 6 ? (int)layer_num - 1 : layer++) << std::endl;
            output << op_title(r.op) << std::endl;
            output << "time : " << (long)r.milliseconds << " ms" << std::endl;
            output << "remaining level : " << r.remaining_level << std::endl;
            output << "scale: " << r.scale << std::endl << std::endl;
        }
        std::size_t label = 0;
        double max_score = -100.0;
        output << "( ";
        for (std::size_t i = 0; i < logits.size(); i++)
        {
            output << "(" << logits[i] << ",0)" << (i + 1 < logits.size() ? ", " : ")");
            if (max_score < logits[i])
            {
                label = i;
                max_score = logits[i];
            }
        }
        output << std::endl;
        output << "total time : " << (long)total_ms << " ms" << std::endl;
        output << "image label: " << image_label << std::endl;
        output << "inferred label: " << label << std::endl;
        output << "max score: " << max_score << std::endl;
        std::lock_guard<std::mutex> guard(report_mu);
        out_share << "image_id: " << image_id << ", image label: " << image_label << ", inferred label: " << label << std::endl;
        std::cout << "image " << image_id << ": total time : " << (long)total_ms << " ms, inferred label: " << label
                  << ", max score: " << max_score << std::endl;
    }
    }
    catch (...)
    {
        std::lock_guard<std::mutex> guard(report_mu);
        if (!failure)
            failure = std::current_exception();
    }
    };
    if (in_flight <= 1)
        image_loop();
    else
    {
        vector<std::thread> threads;
        for (int t = 0; t < in_flight; t++)
            threads.emplace_back(image_loop);
        for (auto &t : threads)
            t.join();
    }
    if (failure)
        std::rethrow_exception(failure);
    const double all_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - all_start).count();
    std::cout << "all threads time : " << (long)all_ms << " ms" << std::endl;
    out_share << std::endl << "all threads time : " << (long)all_ms << " ms" << std::endl;
}
