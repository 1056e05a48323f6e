"""oracle/resnet_ref.py — TEST INFRASTRUCTURE (see oracle/__init__.py).

Functional fp32 CPU restatement of the reference's ResNet18-GroupNorm vision encoder: `VisionEncoder()` of
models/Unet_FiLmLayer.py:316-386 = torchvision.models.resnet18 (BasicBlock x [2, 2, 2, 2]) with `fc = Identity` and every
BatchNorm2d replaced by GroupNorm(C // 16, C).  Driven by a plain state_dict with torchvision's key names.  Pinned against the
reference's own module by oracle/make_golden.py -> tests/golden/resnet18gn.npz.
"""
import torch
import torch.nn.functional as F

BLOCKS = (("layer1.0", 64, 64, 1), ("layer1.1", 64, 64, 1), ("layer2.0", 64, 128, 2), ("layer2.1", 128, 128, 1),
          ("layer3.0", 128, 256, 2), ("layer3.1", 256, 256, 1), ("layer4.0", 256, 512, 2), ("layer4.1", 512, 512, 1))


def shapes():
    s = {"conv1.weight": (64, 3, 7, 7), "bn1.weight": (64,), "bn1.bias": (64,)}
    for name, cin, cout, stride in BLOCKS:
        s[name + ".conv1.weight"] = (cout, cin, 3, 3)
        s[name + ".bn1.weight"] = (cout,)
        s[name + ".bn1.bias"] = (cout,)
        s[name + ".conv2.weight"] = (cout, cout, 3, 3)
        s[name + ".bn2.weight"] = (cout,)
        s[name + ".bn2.bias"] = (cout,)
        if stride != 1 or cin != cout:
            s[name + ".downsample.0.weight"] = (cout, cin, 1, 1)
            s[name + ".downsample.1.weight"] = (cout,)
            s[name + ".downsample.1.bias"] = (cout,)
    return s


def _gn(sd, p, x):
    w = sd[p + ".weight"]
    return F.group_norm(x, w.numel() // 16, w, sd[p + ".bias"])       # replace_bn_with_gn: features_per_group = 16 (:368-380)


def basic_block(sd, p, x, stride):
    """torchvision BasicBlock.forward: relu(bn2(conv2(relu(bn1(conv1(x))))) + downsample(x))."""
    out = F.relu(_gn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1)))
    out = _gn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], padding=1))
    identity = x
    if p + ".downsample.0.weight" in sd:
        identity = _gn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + identity)


def encode(sd, img):
    """torchvision ResNet._forward_impl with fc = Identity (models/Unet_FiLmLayer.py:317-330).  img (N,3,H,W) -> (N,512)."""
    x = F.conv2d(img, sd["conv1.weight"], stride=2, padding=3)
    x = F.relu(_gn(sd, "bn1", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for name, cin, cout, stride in BLOCKS:
        x = basic_block(sd, name, x, stride)
    return torch.flatten(F.adaptive_avg_pool2d(x, (1, 1)), 1)
