"""Names kept from reference constants.py:1-4: the 16 cross-attention layers in UNet execution order."""
UNET_LAYERS = [
    "IN01", "IN02", "IN04", "IN05", "IN07", "IN08", "MID", "OUT03", "OUT04",
    "OUT05", "OUT06", "OUT07", "OUT08", "OUT09", "OUT10", "OUT11",
]
