"""Host-only kernel selection of vs_conv_forward (no GPU needed: vs_conv_forward_path / vs_conv_forward_variant make
no launch).  SURVEY H6: a sample's result must not depend on what else is in the batch, so the kernel chosen for a layer
must be a function of the layer geometry and the storage dtype only — never of the batch size."""
import pytest

from spatiotemporal_variable_separation_b200 import _lib as L

BF16, F32 = 1, 0


def geom(dtype, N, H, W, C, K, R, stride, pad, groups=1):
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    return L.Geom(dtype, N, H, W, C, P, Q, K, R, R, stride, pad, groups, 0, 0)


def selection(dtype, N, *layer):
    lib = L.load()
    g = geom(dtype, N, *layer)
    return tuple((lib.vs_conv_forward_path(g, mode), lib.vs_conv_forward_variant(g, mode)) for mode in (L.DIRECT, L.TRANSPOSED))


# (H, W, C, K, R, stride, pad): the conv geometries of the five configurations, bf16 storage
LAYERS = [
    (64, 64, 1, 64, 4, 2, 1), (64, 64, 5, 64, 4, 2, 1),          # first encoder layer / last decoder layer (thin side)
    (32, 32, 64, 128, 4, 2, 1), (16, 16, 128, 256, 4, 2, 1), (8, 8, 256, 512, 4, 2, 1),      # DCGAN body
    (4, 4, 512, 128, 4, 1, 0), (4, 4, 512, 20, 4, 1, 0),          # encoder head / first up-convolution
    (32, 32, 64, 64, 3, 1, 1), (16, 16, 128, 128, 3, 1, 1), (8, 8, 256, 256, 3, 1, 1),        # VGG / ResNet bodies
    (1, 1, 1200, 1200, 1, 1, 0), (1, 1, 20480, 1200, 1, 1, 0),    # WaveEq MLP
    (64, 64, 15, 64, 5, 2, 3), (17, 17, 64, 128, 3, 2, 1), (17, 17, 64, 128, 1, 2, 0),        # ResNet18 stem / stride-2 blocks
    (12, 12, 72, 96, 3, 1, 1), (8, 8, 40, 200, 4, 2, 1),
]


@pytest.mark.parametrize('layer', LAYERS)
@pytest.mark.parametrize('dtype', [BF16, F32])
def test_selection_does_not_depend_on_the_batch_size(layer, dtype):
    want = selection(dtype, 1, *layer)
    for N in (2, 3, 8, 37, 128, 256, 1664, 4000):
        assert selection(dtype, N, *layer) == want, (layer, N)


def test_variants_of_the_benchmarked_layers():
    """mnist DCGAN (nf = 64), bf16: which tap-GEMM kernel takes which layer."""
    sel = lambda *layer: selection(BF16, 1664, *layer)
    # 64 -> 128 at 32x32 -> 16x16: forward (direct) and the decoder's 128 -> 64 up-convolution (transposed): shifted windows
    assert sel(32, 32, 64, 128, 4, 2, 1) == ((1, 3), (1, 3))
    # 128 -> 256 at 16x16 -> 8x8: 256 output channels forward = CTA pairs; transposed (256 -> 128, 8x8 class grid) per class
    assert sel(16, 16, 128, 256, 4, 2, 1) == ((1, 2), (1, 1))
    # 256 -> 512 at 8x8 -> 4x4: CTA pairs both ways (transposed: 512 -> 256)
    assert sel(8, 8, 256, 512, 4, 2, 1) == ((1, 2), (1, 2))
    # thin layers never reach the tap GEMM: im2col tile / col2im / thin streaming kernels
    for path, variant in sel(64, 64, 1, 64, 4, 2, 1):
        assert path != 1 and variant == 0
    # fp32 storage (parity mode): CUDA-core kernels only
    for path, variant in selection(F32, 128, 32, 32, 64, 128, 4, 2, 1):
        assert path in (0, 2) and variant == 0
