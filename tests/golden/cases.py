"""Golden-vector cases shared by make_golden.py (reference run) and the tests (oracle / engine)."""
CASES = {
    # WindowTransformer: token grid 4x6 -> one padded 8x8 window; default-style res_out (1.5x)
    "window_72x104_r1p5": dict(model="WindowTransformer", shape=(1, 3, 72, 104), wseed=1, xseed=11,
                               kw=dict(res_out=(108, 156))),
    # batch 2, integer factor, 4x5 tokens
    "window_64x80_x2": dict(model="WindowTransformer", shape=(2, 3, 64, 80), wseed=2, xseed=12,
                            kw=dict(upscale_factor=2)),
    # odd input: ceil stride-2, floor patch embed, crop before the skip add; 2x3 windows
    "window_131x189_odd": dict(model="WindowTransformer", shape=(1, 3, 131, 189), wseed=3, xseed=13,
                               kw=dict(res_out=(200, 280))),
    "fast_40x56_x2": dict(model="FastTransformer", shape=(1, 3, 40, 56), wseed=4, xseed=14, kw=dict(upscale_factor=2)),
    "fast_40x56_x3": dict(model="FastTransformer", shape=(1, 3, 40, 56), wseed=4, xseed=15, kw=dict(upscale_factor=3)),
    "fast_40x56_x4": dict(model="FastTransformer", shape=(1, 3, 40, 56), wseed=4, xseed=16, kw=dict(upscale_factor=4)),
    "fast_24x32_x6": dict(model="FastTransformer", shape=(1, 3, 24, 32), wseed=5, xseed=17, kw=dict(upscale_factor=6)),
    # H, W not multiples of 8 -> reflect pad before patch embed, crop after unembed
    "fast_36x52_x2_reflect": dict(model="FastTransformer", shape=(2, 3, 36, 52), wseed=6, xseed=18,
                                  kw=dict(upscale_factor=2)),
    # res_out path: factor = ceil(1.5) = 2, then antialiased Resize down to res_out
    "fast_40x56_res60x84": dict(model="FastTransformer", shape=(1, 3, 40, 56), wseed=7, xseed=19,
                                kw=dict(res_out=(60, 84))),
    # ResidualTransformer only accepts 3600 tokens -> 720p input; store every 8th pixel
    "residual_720p_1080p": dict(model="ResidualTransformer", shape=(1, 3, 720, 1280), wseed=8, xseed=20,
                                kw=dict(res_out=(1080, 1920)), stride=8),
}

# natural-image fixtures (tests/golden/make_natural.py): uint8 LR frame, uint8 HR target, the reference's fp32 output as fp16
NATURAL = {
    "natural_window_96x176_r1p5": dict(model="WindowTransformer", wseed=31, kw=dict(res_out=(144, 264))),
    "natural_fast_96x176_x2": dict(model="FastTransformer", wseed=32, kw=dict(upscale_factor=2)),
}
