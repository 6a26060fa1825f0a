// cnn/infer_seal.h - bootstrapped ResNet-20/32/44/56/110 on CIFAR-10 over multiplexed-packed ciphertexts.
//
// Restates ResNet_cifar10_seal_sparse and import_parameters_cifar10 of the reference's
// cnn_ckks/cpu-ckks/single-key/cnn/infer_seal.cpp (:251-584, :3-107): same parameter set (logN 16, primes
// 51 | 46 x 16 | 51 x 14 | 51, Hamming weight 192, scale 2^46), same rotation-key list, three sparse-slot
// bootstrappers (logn 14/13/12), the alpha = 13 minimax ReLU and the same layer schedule.  The reference does all of
// it inside one function that also parses weights and images from text files per image and per OpenMP thread; here
// the set-up (ResNetCifar10 constructor) is separated from the per-image path (infer) so that one set of keys,
// bootstrapping matrices and evaluation trees serves every image, and the per-stage log (operation, remaining level,
// scale, time) is returned instead of being printed.
#pragma once
#include "cnn/cnn_seal.h"
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

struct ResNetParameters
{
    std::vector<std::vector<double>> conv_weight, bn_bias, bn_running_mean, bn_running_var, bn_weight; // [layer_num - 1]
    std::vector<double> linear_weight, linear_bias;                                                    // classes x 64, classes
    // CIFAR-100 variant (the commented-out ResNet_cifar100_seal_sparse, infer_seal.cpp:585-891): the two down-sampling
    // shortcuts are 1x1 stride-2 convolutions with their own batch norm instead of the zero-padding shortcut
    std::vector<std::vector<double>> shortcut_weight, shortcut_bn_bias, shortcut_bn_mean, shortcut_bn_var, shortcut_bn_weight; // [2]
};

// approximation boundary, class count and shortcut type of the two datasets (infer_seal.cpp:253 / :588)
struct ResNetVariant
{
    double B = 40.0;
    int classes = 10;
    bool shortcut_conv = false;
    static ResNetVariant cifar10() { return {}; }
    static ResNetVariant cifar100() { return { 65.0, 100, true }; }
};

// number of residual blocks per stage minus one (infer_seal.cpp:397-402)
int resnet_end_num(std::size_t layer_num);
// (ci, co) of convolution `index` (0 .. layer_num - 2) in the reference's file order
void resnet_conv_shape(std::size_t layer_num, std::size_t index, int &ci, int &co);

// reads <dir>/conv1_weight.txt, layer<j>_<k>_conv<1|2>_weight.txt, *_bn*_{bias,running_mean,running_var,weight}.txt,
// linear_{weight,bias}.txt - the reference's pretrained_parameters/resnet<L>_new layout
void import_parameters_cifar10(std::vector<double> &linear_weight, std::vector<double> &linear_bias,
                               std::vector<std::vector<double>> &conv_weight, std::vector<std::vector<double>> &bn_bias,
                               std::vector<std::vector<double>> &bn_running_mean, std::vector<std::vector<double>> &bn_running_var,
                               std::vector<std::vector<double>> &bn_weight, std::size_t layer_num, std::size_t end_num,
                               const std::string &dir = "../../pretrained_parameters");

// CNN rotation steps of infer_seal.cpp:345-362 (the list is the same for every depth)
const std::vector<int> &resnet_rotation_kinds();

// The reference's entry point (infer_seal.cpp:251-584; run/run_cnn.cpp calls it): parameters, keys, bootstrappers, then
// images start..end one after the other.  Weights from <weights_dir>/resnet<L>_new and images / labels from
// <images_dir>/test_values.txt, test_label.txt when those exist (defaults: the reference's relative paths), otherwise
// random-init weights and synthetic images.  Writes <result_dir>/resnet<L>_cifar10_image<id>.txt in the reference's
// log format and <result_dir>/resnet<L>_cifar10_label_<start>_<end>.
void ResNet_cifar10_seal_sparse(std::size_t layer_num, std::size_t start_image_id, std::size_t end_image_id,
                                const std::string &result_dir = "../../result",
                                const std::string &weights_dir = "../../pretrained_parameters",
                                const std::string &images_dir = "../../../testFile");
// The CIFAR-100 driver the fork has commented out (infer_seal.cpp:585-891; run_cnn.cpp:24): ResNet-32, B = 65, 1x1
// stride-2 shortcut convolutions, 100 classes.  Weights from <weights_dir>/resnet32_cifar100, images from
// <images_dir>/test_values.txt when present, otherwise random-init weights and synthetic images.
void ResNet_cifar100_seal_sparse(std::size_t layer_num, std::size_t start_image_id, std::size_t end_image_id,
                                 const std::string &result_dir = "../../result",
                                 const std::string &weights_dir = "../../pretrained_parameters",
                                 const std::string &images_dir = "../../../testFile_cifar100");
// reads <dir>/... as import_parameters_cifar10 plus layer{2,3}_0_shortcut_0_weight.txt and
// layer{2,3}_0_shortcut_1_{bias,running_mean,running_var,weight}.txt (infer_seal.cpp:108-250)
void import_parameters_cifar100(ResNetParameters &parameters, std::size_t layer_num, const std::string &dir);

struct ResNetTraceRow
{
    int op;               // 0 conv, 1 bn, 2 relu, 3 bootstrap, 4 add, 5 downsample, 6 avgpool, 7 fc
    int remaining_level;  // chain_index of the result
    double scale;
    double milliseconds;
};

class ResNetCifar10
{
public:
    // constants of infer_seal.cpp:253-304
    static constexpr long alpha = 13, comp_no = 3;
    static constexpr double scaled_val = 1.7;
    static constexpr long boundary_K = 25, boot_deg = 59, scale_factor = 2, inverse_deg = 1, logN = 16, loge = 10, logn = 15;
    static constexpr int logp = 46, logq = 51, log_special_prime = 51, remaining_level = 16, boot_level = 14;
    static std::vector<int> coeff_bit_vec();

    ResNetCifar10(std::size_t layer_num, ResNetParameters parameters, seal::SEALContext &context, seal::KeyGenerator &keygen,
                  seal::CKKSEncoder &encoder, seal::Encryptor &encryptor, seal::Decryptor &decryptor, seal::Evaluator &evaluator,
                  seal::PublicKey &public_key, seal::SecretKey &secret_key, seal::RelinKeys &relin_keys, seal::GaloisKeys &gal_keys,
                  ResNetVariant variant = ResNetVariant::cifar10());
    ~ResNetCifar10();

    // every rotation step the network and its bootstrappers use (hand to KeyGenerator::create_galois_keys)
    std::vector<int> galois_steps() const;
    // after the keys exist: LT coefficients of the three bootstrappers
    void prepare();

    // image: 32*32*3 values (channel-major).  Returns the 10 logits; appends one row per operation to `trace`.
    std::vector<double> infer(const std::vector<double> &image, std::vector<ResNetTraceRow> *trace = nullptr);
    // the three phases of infer(): client-side packing and encryption (infer_seal.cpp:434-453), the encrypted network
    // (:455-537), client-side decryption of the ten logits (:543-551)
    TensorCipher encrypt_image(const std::vector<double> &image);
    TensorCipher infer_encrypted(const TensorCipher &input, std::vector<ResNetTraceRow> *trace = nullptr);
    std::vector<double> decrypt_logits(const TensorCipher &output);

    std::size_t layer_num() const { return layer_num_; }
    int classes() const { return variant_.classes; }
    const double B; // approximation boundary: activations are carried divided by B

private:
    std::size_t layer_num_;
    int end_num_;
    ResNetVariant variant_;
    ResNetParameters w_;
    seal::SEALContext &context_;
    seal::KeyGenerator &keygen_;
    seal::CKKSEncoder &encoder_;
    seal::Encryptor &encryptor_;
    seal::Decryptor &decryptor_;
    seal::Evaluator &evaluator_;
    seal::PublicKey &public_key_;
    seal::SecretKey &secret_key_;
    seal::RelinKeys &relin_keys_;
    seal::GaloisKeys &gal_keys_;
    std::vector<int> deg_{ 15, 15, 27 };
    std::vector<minicomp::Tree> tree_;
    Bootstrapper *boot_[3];
    std::vector<std::unique_ptr<ConvPlan>> conv_plans_; // one per convolution, built at its first use
    std::mutex plan_mu_;
    bool prepared_ = false;
};
