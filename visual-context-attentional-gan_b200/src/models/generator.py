"""Drop-in for the reference's src/models/generator.py: `from src.models.generator import Decoder, Discriminator,
gan_loss, sync_Discriminator, Postnet` (train.py:8, test.py:8) resolves to the B200-native modules."""
from vcagan_b200.models import (ResBlk1D, ResBlk, GenResBlk, Flatten, Avgpool, AVAttention, Postnet, Decoder,  # noqa: F401
                                Discriminator, sync_Discriminator, gan_loss, final_length, BasicBlock)
