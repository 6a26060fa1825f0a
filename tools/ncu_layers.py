"""ncu target: one minimax ReLU (17 -> 3 limbs) and one stage-1 convolution (16 -> 16 channels, 32 x 32) at N = 2^16
in hybrid mode between cudaProfilerStart / cudaProfilerStop - the small-kernel phases of a ResNet image.
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off ... python tools/ncu_layers.py"""
import ctypes
import os
import sys

os.environ["B200CKKS_HYBRID_KS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fhe-gpt-2_b200", "python"), os.path.join(ROOT, "oracle")]
import numpy as np
import plain_model as pm
from b200ckks.app import App

cudart = ctypes.CDLL("libcudart.so")
BITS = [51] + [46] * 16 + [51] * 14 + [51]
s = App().session(16, BITS, hamming_weight=192)
from b200ckks import synthetic
net = s.resnet(20, synthetic.random_weights(20, seed=0))  # declares the rotation keys of the network
rng = np.random.default_rng(3)
x = rng.uniform(-0.9, 0.9, s.slots)
t = rng.uniform(-0.25, 0.25, (16, 32, 32))
wt, var, bw = rng.normal(0, 0.1, 9 * 16 * 16), rng.uniform(0.5, 1.5, 16), rng.uniform(0.5, 1.0, 16)
parms = [1, 32, 32, 16, 16, 2, 15]


def relu_in():
    c = s.encrypt(x, 2.0 ** 46)
    s.mod_switch_to(c, 17)
    return c


def conv_in():
    c = s.encrypt(pm.pack(t, 1, 2, s.slots), 2.0 ** 46)
    s.mod_switch_to(c, 3)
    return c


for _ in range(2):
    s.relu(relu_in())
    s.conv(conv_in(), parms, 16, 1, wt, var, bw)
a, b = relu_in(), conv_in()
s.sync()
which = sys.argv[1] if len(sys.argv) > 1 else "both"
cudart.cudaProfilerStart()
if which in ("both", "relu"):
    s.relu(a)
if which in ("both", "conv"):
    s.conv(b, parms, 16, 1, wt, var, bw)
s.sync()
cudart.cudaProfilerStop()
s.close()
