"""Pin the read-level featuriser oracle (oracle/read_matrix_oracle.py, calculate_read_alignment restated) on the
reference's own numbers (build container only).

Run:  python tests/golden/make_read_matrix_golden.py     (needs /root/reference; writes tests/golden/read_matrix.npz)

1. The reference's literal expectations for its four mock reads (medaka/test/mock_data.py:22-100 ->
   medaka/test/test_read_alignment_matrix.py:103-180 without dwells, :182-259 with dwells: move tables, a missing
   quality string, a malformed move table) are parsed out of the test source with `ast` and stored next to the reads.
2. The real-BAM regression numbers of the same test file (:27-71, produced by the htslib-based C): feature matrix
   (86294, 45, 4) for utg000001l:50000-100000, (86294, 291, 4) with row_per_read / max_reads=1000, positions equal to
   the counts featuriser's, reads per column and base counts per column equal to the sym_indels counts - ASSERTED here
   on the oracle's output, which pins the restated row bookkeeping on real data.
3. A slice (utg000001l:50000-50250) with the oracle's matrix is stored as the fixture the CPU and GPU tests replay.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from medaka_b200 import bam  # noqa: E402
from oracle import features_oracle, pileup_oracle, read_matrix_oracle  # noqa: E402

REF = "/root/reference/medaka/test"
BAM = os.path.join(REF, "data", "test_reads.bam")


def mock_reads():
    """simple_data['calls'] as oracle records (array.array / list arithmetic in the literal evaluated by hand)."""
    src = open(os.path.join(REF, "mock_data.py")).read()
    tree = ast.parse(src)
    node = [n for n in tree.body if isinstance(n, ast.Assign) and n.targets[0].id == "simple_data"][0]
    import array
    data = eval(compile(ast.Expression(node.value), "mock_data", "eval"), {"array": array})
    recs = []
    for c in data["calls"]:
        recs.append(dict(query_name=c["query_name"], pos=0, cigar=c["cigarstring"], seq=c["seq"], flag=c["flag"],
                         mapq=c.get("mapping_quality", 60), qual=list(c["quality"]),
                         tags={k: v for k, v in c["tags"].items() if k in ("DT", "mv", "HP", "AA")}))
    return recs


def expected_arrays():
    """The two `features=np.array([...], dtype='int8').swapaxes(0, 1)` literals of test_030 / test_031."""
    src = open(os.path.join(REF, "test_read_alignment_matrix.py")).read()
    tree = ast.parse(src)
    out = {}
    for fn in ast.walk(tree):
        if isinstance(fn, ast.FunctionDef) and fn.name in ("test_030_bams_to_training_samples_simple",
                                                           "test_031_bams_to_training_samples_dwells"):
            for call in ast.walk(fn):
                if isinstance(call, ast.keyword) and call.arg == "features":
                    lit = call.value.func.value.args[0]          # np.array(<list>, dtype=..).swapaxes(0, 1)
                    out[fn.name] = np.array(ast.literal_eval(lit), dtype=np.int8).swapaxes(0, 1)
    return out["test_030_bams_to_training_samples_simple"], out["test_031_bams_to_training_samples_dwells"]


def batch_records(rb):
    recs = pileup_oracle.records_from_batch(rb)
    qoff = np.concatenate([[0], np.cumsum(rb.l_seq)])
    for i, r in enumerate(recs):
        r["query_name"] = rb.names[i]
        r["qual"] = rb.qual[qoff[i]:qoff[i + 1]].tolist()
        r["tags"] = rb.tags[i]
    return recs


def main():
    recs = mock_reads()
    exp_plain, exp_dwell = expected_arrays()
    r0 = [dict(r) for r in recs]
    r0[0]["qual"] = None                                          # "we had a bug caused by missing qualities" (:107-108)
    m, pos, _, _ = read_matrix_oracle.read_alignment(r0, 0, 100)
    assert m.shape == (9, 4, 4) and np.array_equal(m, exp_plain), "mock reads, no dwells"
    md, posd, _, _ = read_matrix_oracle.read_alignment(r0, 0, 100, include_dwells=True)
    assert md.shape == (9, 4, 5) and np.array_equal(md, exp_dwell), "mock reads, dwells"
    print("mock-read literals of test_read_alignment_matrix.py reproduced (with and without dwells)")

    bf = bam.BamFile(BAM)
    rb = bf.fetch("utg000001l", 50000, 100000, with_names=True, with_qual=True, with_tags=True)
    big = batch_records(rb)
    mat, p, _, _ = read_matrix_oracle.read_alignment(big, 50000, 100000)
    assert mat.shape == (86294, 45, 4) and tuple(p[0]) == (50000, 0) and tuple(p[-1]) == (99999, 1)
    mat_rpr, _, _, _ = read_matrix_oracle.read_alignment(big, 50000, 100000, row_per_read=True, max_reads=1000)
    assert mat_rpr.shape == (86294, 291, 4)
    counts, cpos = pileup_oracle.pileup_counts_from_batch(rb, 50000, 100000)
    assert np.array_equal(p, cpos)
    sym, _ = features_oracle.post_process_pileup(counts.copy(), cpos, None, sym_indels=True)     # test_002 (:41-60)
    assert np.array_equal((mat[:, :, 0] != 0).sum(-1), sym.sum(-1))
    base_counts = np.array([(mat[:, :, 0] == (i + 1)).sum(-1) for i in range(5)])
    want = np.hstack([sym[:, :4] + sym[:, 4:8], sym[:, 8][:, None] + sym[:, 9][:, None]]).transpose()
    assert np.array_equal(base_counts, want)
    print("real-BAM regression numbers reproduced: (86294, 45, 4), row_per_read (86294, 291, 4), totals == sym_indels counts")

    start, end = 50000, 50250
    rb = bf.fetch("utg000001l", start, end, with_names=True, with_qual=True, with_tags=True, with_aux=True)
    small = batch_records(rb)
    out = {}
    for key, kw in (("default", {}), ("rpr", dict(row_per_read=True, max_reads=1000)), ("max5", dict(max_reads=5)),
                    ("hap", dict(include_haplotype=True))):
        mm, pp, left, right = read_matrix_oracle.read_alignment(small, start, end, **kw)
        out["mat_" + key] = mm
        out["left_" + key] = np.array(left)
        out["right_" + key] = np.array(right)
    # 4. sub-region results of the same slice joined by the REFERENCE's own `__enforce_read_matrix_chunk_contiguity`
    #    (medaka/features.py:470-557, imported unmodified behind the stand-ins of make_golden.py): the row re-ordering on
    #    read identity and the padding that medaka_b200.features._join_read_matrix_chunks restates.
    import make_golden
    make_golden.install_stubs()
    sys.path.insert(0, "/root/reference")
    import medaka.features as ref_features
    join = getattr(ref_features, "__enforce_read_matrix_chunk_contiguity")
    cuts = [(50000, 50090), (50090, 50170), (50170, 50250)]
    chunk_in = []
    for k, (a, b) in enumerate(cuts):
        mm, pp_, left, right = read_matrix_oracle.read_alignment(small, a, b)
        out["chunk%d_mat" % k] = mm
        out["chunk%d_major" % k] = pp_["major"]
        out["chunk%d_minor" % k] = pp_["minor"]
        out["chunk%d_left" % k] = np.array([x.encode() for x in left], dtype="S")
        out["chunk%d_right" % k] = np.array([x.encode() for x in right], dtype="S")
        chunk_in.append((mm, pp_, (out["chunk%d_left" % k], out["chunk%d_right" % k])))
    joined = join(chunk_in)
    assert len(joined) == 1
    out["joined_mat"] = joined[0][0]
    out["joined_major"] = joined[0][1]["major"]
    print("reference chunk join:", [c[0].shape for c in chunk_in], "->", joined[0][0].shape)
    # 5. a harder join: short synthetic reads that come and go, five sub-regions of unequal depth, one of them behind a
    #    coverage gap (so: row re-ordering, rows filled from vacated slots, appended rows, padding, a split)
    from oracle import synth
    rs = np.random.RandomState(12)
    srecs = synth.synth_reads(160, 2600, seed=77, mean_len=260)
    srecs = [r for r in srecs if not (1480 <= r["pos"] < 1640)]
    srecs.sort(key=lambda r: r["pos"])
    for i, r in enumerate(srecs):
        r["query_name"], r["tags"] = "s%d" % i, {}
        r["qual"] = rs.randint(1, 40, len(r["seq"])).tolist()
    import re
    for r in srecs:                                   # nothing may bridge the gap
        if r["pos"] < 1480:
            used, out_ops = 0, []
            for n_, op in re.findall(r"(\d+)([MIDNSHP=X])", r["cigar"]):
                n_ = int(n_)
                if op in "MDN=X":
                    if r["pos"] + used + n_ > 1480:
                        n_ = 1480 - r["pos"] - used
                        if n_ > 0 and op in "M=X":
                            out_ops.append("%d%s" % (n_, op))
                        break
                    used += n_
                out_ops.append("%d%s" % (n_, op))
            while out_ops and out_ops[-1][-1] not in "M=X":
                out_ops.pop()
            r["cigar"] = "".join(out_ops) or "1M"
    hard_in = []
    for k, (a, b) in enumerate([(100, 600), (600, 1100), (1100, 1700), (1700, 2100), (2100, 2500)]):
        mm, pp_, left, right = read_matrix_oracle.read_alignment(srecs, a, b)
        out["hard%d_mat" % k] = mm
        out["hard%d_major" % k] = pp_["major"]
        out["hard%d_minor" % k] = pp_["minor"]
        out["hard%d_left" % k] = np.array([x.encode() for x in left], dtype="S")
        out["hard%d_right" % k] = np.array([x.encode() for x in right], dtype="S")
        hard_in.append((mm, pp_, (out["hard%d_left" % k], out["hard%d_right" % k])))
    hard = join(hard_in)
    out["hard_n"] = len(hard)
    for i, (mm, pp_) in enumerate(hard):
        out["hard_joined%d_mat" % i] = mm
        out["hard_joined%d_major" % i] = pp_["major"]
    print("reference chunk join (hard):", [c[0].shape for c in hard_in], "->", [h[0].shape for h in hard])
    name_off = np.concatenate([[0], np.cumsum([len(n) for n in rb.names])]).astype(np.int64)
    np.savez_compressed(
        os.path.join(HERE, "read_matrix.npz"),
        meta="mock literals of test_read_alignment_matrix.py:103-259; test_reads.bam utg000001l:%d-%d, %d records" % (
            start, end, len(rb.pos)),
        mock_plain=exp_plain, mock_dwell=exp_dwell,
        start=start, end=end, pos=rb.pos, flag=rb.flag, mapq=rb.mapq, cigar=rb.cigar, cigar_off=rb.cigar_off,
        seq=rb.seq, seq_off=rb.seq_off, l_seq=rb.l_seq, qual=rb.qual, names="\n".join(rb.names), name_off=name_off,
        aux=rb.aux, aux_off=rb.aux_off,
        major=pp["major"], minor=pp["minor"], **out)
    print("fixture:", len(rb.pos), "records,", out["mat_default"].shape)


if __name__ == "__main__":
    main()
