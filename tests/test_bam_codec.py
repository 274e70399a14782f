"""CPU: the test-side BAM record encoder and the restatement of parseBamAlignment (sam/bam-files.go:314-400) are inverse
to each other -- they are what the device ingest (elp_append_bam) is compared against in test_gpu_parity.py."""
import numpy as np

from elprep_b200 import sam, synth
from util import decode_bam, encode_bam


def test_bam_roundtrip():
    w = synth.make_workload(1500, [("chr20", 200_000), ("chr21", 100_000)], seed=3, want_reference=False, unmapped_frac=0.1)
    raw, offs = encode_bam(w.batch, w.header)
    b = decode_bam(raw, offs, w.header)
    for f in sam.AlignmentBatch.FIELDS:
        assert np.array_equal(getattr(b, f), getattr(w.batch, f)), f
    # fixed-field offsets of the format (sam/bam-files.go:300-312): refID at 4, l_seq at 20, name at 36 of a record with its block_size
    r0 = raw[:int(offs[1])].tobytes()
    assert int.from_bytes(r0[4:8], "little", signed=True) == int(w.batch.refid[0]) and int.from_bytes(r0[20:24], "little") == int(w.batch.lseq[0])
    assert r0[36:36 + r0[12] - 1] == bytes(w.batch.qname[:int(w.batch.qname_off[1])])
