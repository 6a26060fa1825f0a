// run/run_cnn.cpp - `./cnn <layers> <dataset> <start image> <end image>`: the reference's CLI
// (cnn_ckks/run/run_cnn.cpp:8-27) on the engine.
//
//   ./cnn 20 10 0 0      bootstrapped ResNet-20 on CIFAR-10 image 0
//
// Weights are read from ../../pretrained_parameters/resnet<L>_new (the reference's layout) when that directory
// exists, images and labels from ../../../testFile/; otherwise random-init weights and synthetic images are used and
// the label check is skipped.  One log per image is written in the reference's format
// (result/resnet<L>_cifar10_image<id>.txt: operation, time, remaining level, scale per stage; logits; label).
#include "cnn/infer_seal.h"
#include <cstdlib>
#include <filesystem>
#include <iostream>

int main(int argc, char **argv)
{
    if (argc < 5)
    {
        std::cerr << "usage: " << argv[0] << " <layers 20|32|44|56|110> <dataset 10> <start image> <end image> [result dir]\n";
        return 2;
    }
    const int layer = std::atoi(argv[1]), dataset = std::atoi(argv[2]), start = std::atoi(argv[3]), end = std::atoi(argv[4]);
    if (start < 0 || start >= 10000)
        throw std::invalid_argument("start number is not correct");
    if (end < 0 || end >= 10000)
        throw std::invalid_argument("end number is not correct");
    if (start > end)
        throw std::invalid_argument("start number is larger than end number");
    std::cout << "model: ResNet-" << layer << std::endl;
    std::cout << "dataset: CIFAR-" << dataset << std::endl;
    std::cout << "start image: " << start << std::endl;
    std::cout << "end image: " << end << std::endl;
    if (dataset != 10 && dataset != 100)
        throw std::invalid_argument("dataset number is not correct");
    const std::string result_dir = argc > 5 ? argv[5] : "result";
    std::filesystem::create_directories(result_dir);
    // run_cnn.cpp:23-25: the CIFAR-100 call is commented out in the fork; it is live here
    if (dataset == 10)
        ResNet_cifar10_seal_sparse((std::size_t)layer, (std::size_t)start, (std::size_t)end, result_dir);
    else
        ResNet_cifar100_seal_sparse((std::size_t)layer, (std::size_t)start, (std::size_t)end, result_dir);
    return 0;
}
