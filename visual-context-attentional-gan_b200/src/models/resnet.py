"""Drop-in for the reference's src/models/resnet.py (same import path and names); implementation in
vcagan_b200.models (hand-written sm_100a kernels behind the reference module API)."""
from vcagan_b200.models import (conv3x3, downsample_basic_block, downsample_basic_block_v2, BasicBlock, ResNet)  # noqa: F401
