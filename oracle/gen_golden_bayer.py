"""Golden vectors of the UNMODIFIED reference LfpDecoder.comp_bayer (through oracle/ref_shim.py).
Build container only:  python oracle/gen_golden_bayer.py   # writes tests/golden/comp_bayer.npz
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref_shim.load_reference()
    from plenopticam.lfp_reader.lfp_decoder import LfpDecoder
    rng = np.random.default_rng(51)
    out = {}
    for bits, shape in ((10, (48, 36)), (12, (40, 34))):         # shape = (width, height) like the Lytro metadata
        n = shape[0] * shape[1]
        nbytes = n * bits // 8
        buf = rng.integers(0, 256, size=nbytes, dtype=np.uint8).tobytes()
        with tempfile.TemporaryDirectory() as tmp:
            cfg, sta = ref_shim.make_cfg_sta(tmp, onp.synth_centroids(3, 3, 9.0, 'rec', jitter=0, seed=0), 'rec', 7)
            cfg.lfpimg = {'bit': bits}
            dec = LfpDecoder(cfg=cfg, sta=sta, lfp_path=os.path.join(tmp, 'x.lfr'))
            dec._img_buf, dec._shape = bytearray(buf), list(shape)
            assert dec.comp_bayer() is True
            ref = np.asarray(dec._bay_img)
        mine = onp.comp_bayer(buf, shape, bits)
        print(bits, ref.shape, ref.dtype, "oracle == reference:", np.array_equal(mine, ref))
        out["buf%d" % bits] = np.frombuffer(buf, dtype=np.uint8)
        out["shape%d" % bits] = np.array(shape)
        out["img%d" % bits] = ref
    np.savez_compressed(os.path.join(GOLDEN, "comp_bayer.npz"), **out)


if __name__ == "__main__":
    main()
