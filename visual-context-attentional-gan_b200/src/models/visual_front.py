"""Drop-in for the reference's src/models/visual_front.py: `from src.models.visual_front import Visual_front`
(train.py:7, test.py:7) resolves to the B200-native module."""
from vcagan_b200.models import Visual_front, BasicBlock, ResNet  # noqa: F401
