"""The reference's two command-line drivers on the engine (cnn_ckks/run/run_cnn.cpp, run_bootstrapping.cpp):
`cnn <layers> <dataset> <start> <end>` and `run_bootstrapping`, run as a user of the reference would."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fhe-gpt-2_b200", "lib")


def _run(args, timeout, cwd=None):
    return subprocess.run(args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)


def test_cli_usage_and_argument_checks():
    # no GPU work before the arguments are validated (run_cnn.cpp:13-25 of the reference throws on the same cases)
    exe = os.path.join(LIB, "cnn")
    assert os.access(exe, os.X_OK), "build() must produce lib/cnn"
    assert os.access(os.path.join(LIB, "run_bootstrapping"), os.X_OK)
    r = _run([exe], 60)
    assert r.returncode == 2 and "usage" in r.stdout
    r = _run([exe, "20", "10", "5", "3"], 60)
    assert r.returncode != 0 and "start number is larger than end number" in r.stdout


@pytest.mark.gpu
def test_run_bootstrapping_cli():
    r = _run([os.path.join(LIB, "run_bootstrapping")], 900)
    assert r.returncode == 0, r.stdout[-2000:]
    m = re.search(r"Absolute mean of error: ([0-9.eE+-]+)", r.stdout)
    assert m, r.stdout[-2000:]
    assert float(m.group(1)) < 1e-4  # the reference prints ~1e-6..1e-5 for this parameter set


@pytest.mark.gpu
def test_cnn_cli_resnet20_log_format(tmp_path):
    out = str(tmp_path / "result")
    r = _run([os.path.join(LIB, "cnn"), "20", "10", "0", "0", out], 1500)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "model: ResNet-20" in r.stdout and "dataset: CIFAR-10" in r.stdout
    log = open(os.path.join(out, "resnet20_cifar10_image0.txt")).read()
    # one "layer k" section per convolution layer + the pooling/FC section, as in the reference's result logs
    assert [int(x) for x in re.findall(r"^layer (\d+)$", log, re.M)][:3] == [0, 1, 2]
    assert log.count("approximate ReLU...") == 19
    assert log.count("bootstrapping...") == 18  # the first ReLU still has levels from encryption
    assert log.count("multiplexed parallel downsampling...") == 2
    assert log.count("average pooling...") == 1 and log.count("fully connected layer...") == 1
    levels = [int(x) for x in re.findall(r"remaining level : (\d+)", log)]
    assert levels and max(levels) <= 31 and min(levels) >= 0
    logits = [float(x) for x in re.findall(r"\(([-0-9.eE+]+),0\)", log)]
    assert len(logits) == 10 and all(abs(v) < 40.0 for v in logits)
    inferred = int(re.search(r"inferred label: (\d+)", log).group(1))
    assert inferred == max(range(10), key=lambda i: logits[i])
    assert re.search(r"total time : \d+ ms", log)
    share = open(os.path.join(out, "resnet20_cifar10_label_0_0")).read()
    assert "image_id: 0" in share and "all threads time" in share


@pytest.mark.gpu
def test_cnn_cli_images_in_flight(tmp_path):
    """./cnn 20 10 0 2 with two images in flight (the reference's OpenMP image loop): one log per image, same logits
    for the same synthetic image as a run with one image at a time would give (seeded keys)."""
    out = str(tmp_path / "result")
    env = dict(os.environ, B200CKKS_IMAGES_IN_FLIGHT="2", B200CKKS_SEED="0x5EA1C0DE")
    r = subprocess.run([os.path.join(LIB, "cnn"), "20", "10", "0", "2", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=1500, env=env)
    assert r.returncode == 0, r.stdout[-2000:]
    logits = []
    for i in range(3):
        log = open(os.path.join(out, f"resnet20_cifar10_image{i}.txt")).read()
        logits.append([float(x) for x in re.findall(r"\(([-0-9.eE+]+),0\)", log)])
        assert len(logits[-1]) == 10 and log.count("bootstrapping...") == 18
    assert logits[0] != logits[1]                      # different synthetic images
    share = open(os.path.join(out, "resnet20_cifar10_label_0_2")).read()
    assert sorted(int(x) for x in re.findall(r"image_id: (\d+)", share)) == [0, 1, 2]
