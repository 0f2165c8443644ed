"""melogan: host side of the B200-native Melo-GAN hot path (ctypes over the C-ABI in include/melogan_b200.h)."""
