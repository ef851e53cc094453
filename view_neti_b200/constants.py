"""Names kept from reference constants.py:1-4: the 16 cross-attention layers in UNet execution order."""
UNET_LAYERS = [
    "IN01", "IN02", "IN04", "IN05", "IN07", "IN08", "MID", "OUT03", "OUT04",
    "OUT05", "OUT06", "OUT07", "OUT08", "OUT09", "OUT10", "OUT11",
]

# reference constants.py:6-11: the 50 DPM/DDIM inference timesteps the prompt embeddings are precomputed for
SD_INFERENCE_TIMESTEPS = [999 - 20 * i if i < 25 else 500 - 20 * (i - 25) for i in range(50)]
