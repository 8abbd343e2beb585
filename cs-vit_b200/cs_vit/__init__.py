"""cs_vit - B200-native drop-in for the CS-ViT hot path."""
__version__ = "0.1.0"
