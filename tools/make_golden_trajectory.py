"""Parses the reference's committed run log (result/resnet20_cifar10_image0.txt) into tests/golden/
resnet20_trajectory.json: operation, remaining level and scale per stage, per-operation time, logits, label."""
import json
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/result/resnet20_cifar10_image0.txt"
names = {"multiplexed parallel convolution...": "conv", "multiplexed parallel batch normalization...": "bn",
         "approximate ReLU...": "relu", "bootstrapping...": "bootstrap", "cipher add...": "add",
         "multiplexed parallel downsampling...": "downsample", "average pooling...": "avgpool", "fully connected layer...": "fc"}
rows, cur, total, logits, label = [], None, None, None, None
for line in open(src):
    line = line.strip()
    if line in names:
        cur = {"op": names[line], "ms": None, "level": None, "scale": None}
        rows.append(cur)
    elif line.startswith("time :") and cur is not None:
        cur["ms"] = float(line.split()[2])
    elif line.startswith("remaining level :") and cur is not None:
        cur["level"] = int(line.split()[-1])
    elif line.startswith("scale:") and cur is not None:
        cur["scale"] = float(line.split()[-1])
    elif line.startswith("total time"):
        total = float(line.split()[3])
    elif line.startswith("inferred label"):
        label = int(line.split()[-1])
    elif line.startswith("( (") and "layer" not in line and len(re.findall(r"\(", line)) == 11:
        logits = [float(m) for m in re.findall(r"\((-?[0-9.e+-]+),", line)]
json.dump({"source": "result/resnet20_cifar10_image0.txt of the reference", "rows": rows, "total_ms": total, "logits": logits,
           "inferred_label": label}, open("tests/golden/resnet20_trajectory.json", "w"), indent=0)
print(len(rows), total, label, logits)
