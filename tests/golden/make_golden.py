"""Regenerates the fixtures in tests/golden/ from the reference's data files.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the committed fixtures.

  golden_input_chirp_1024.npy   first 1024 `I Q` lines of
                                /root/reference/data/golden_input_chirp.txt (int16 [1024, 2]);
                                the file is a single complex tone, 0.283333 cycles/sample.
  radar_output_summary.npz      structural summary of /root/reference/data/radar_output.txt
                                (131072 lines `range doppler 0 0 magnitude`, range-major):
                                per-row max / sum, the six target rows in full, global stats.
                                SURVEY.md Appendix B explains why only structural checks
                                (not bit-exact parity) are possible against this file.
"""
import os
import numpy as np

REF = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = np.loadtxt(os.path.join(REF, "golden_input_chirp.txt"), dtype=np.int64)
    assert g.shape == (2000, 2)
    np.save(os.path.join(HERE, "golden_input_chirp_1024.npy"), g[:1024].astype(np.int16))

    r = np.loadtxt(os.path.join(REF, "radar_output.txt"), dtype=np.int64)
    assert r.shape == (131072, 5)
    rng_idx, dop_idx, mag = r[:, 0], r[:, 1], r[:, 4]
    assert np.array_equal(rng_idx, np.repeat(np.arange(1024), 128))      # range-major
    assert np.array_equal(dop_idx, np.tile(np.arange(128), 1024))        # Doppler fast
    assert not r[:, 2].any() and not r[:, 3].any()
    m = mag.reshape(1024, 128)
    rows = np.array([99, 100, 101, 499, 500, 501])
    np.savez_compressed(
        os.path.join(HERE, "radar_output_summary.npz"),
        row_max=m.max(axis=1).astype(np.int32), row_sum=m.sum(axis=1).astype(np.int64),
        target_rows=rows, target_row_values=m[rows].astype(np.int32),
        median=np.int64(np.median(m)), maximum=np.int64(m.max()),
        argmax=np.array(np.unravel_index(m.argmax(), m.shape)), n_zero=np.int64((m == 0).sum()))


if __name__ == "__main__":
    main()
