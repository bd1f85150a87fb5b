from geniconet_b200.ico_conv import IcoConvS2S, IcoUpsampleS2S  # noqa: F401
