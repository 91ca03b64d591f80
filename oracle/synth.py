"""Deterministic synthetic weights / pileup inputs shared by the oracle, tests and bench.

TEST INFRASTRUCTURE (see oracle/__init__.py): nothing under ``medaka_b200/`` imports this.

Everything here uses ``numpy.random.RandomState`` (frozen bit-stream guarantee) so
that fixtures made in the build container (tests/golden/make_golden.py, which runs
the real reference classes) can be regenerated bit-for-bit on the GPU box where
/root/reference does not exist.

State-dict key names and shapes follow torch.nn.GRU / Linear exactly as the
reference's GRUModel uses them (medaka/architectures/gru.py:46-55): gate order
r,z,n; ``gru.weight_ih_l{k}[_reverse]`` [3H,in], ``gru.weight_hh_l{k}[_reverse]``
[3H,H], ``gru.bias_ih/bias_hh`` [3H], ``linear.weight`` [5,2H], ``linear.bias`` [5].
"""
import numpy as np

GATES = 3


def state_dict_keys(n_layers=2, bidirectional=True):
    keys = []
    for layer in range(n_layers):
        for sfx in ([""] + (["_reverse"] if bidirectional else [])):
            for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                keys.append("gru.{}_l{}{}".format(name, layer, sfx))
    keys += ["linear.weight", "linear.bias"]
    return keys


def synth_state_dict(seed=0, num_features=10, gru_size=128, n_layers=2,
                     bidirectional=True, head_gain=8.0, rec_gain=1.0):
    """Seeded float32 weights, torch-default scale U(-1/sqrt(H), 1/sqrt(H)).

    ``head_gain`` widens the Linear so logits have a realistic spread (random
    default-init heads give near-uniform softmax, i.e. nothing but near-ties;
    SURVEY.md section 7, "Precision vs parity").
    Returns {key: np.ndarray(float32)} with torch's state-dict key names.
    """
    rs = np.random.RandomState(seed)
    H = gru_size
    ndir = 2 if bidirectional else 1
    bound = 1.0 / np.sqrt(H)
    sd = {}
    for layer in range(n_layers):
        n_in = num_features if layer == 0 else H * ndir
        for sfx in ([""] + (["_reverse"] if bidirectional else [])):
            sd["gru.weight_ih_l{}{}".format(layer, sfx)] = rs.uniform(
                -bound, bound, (GATES * H, n_in)).astype(np.float32)
            sd["gru.weight_hh_l{}{}".format(layer, sfx)] = (rec_gain * rs.uniform(
                -bound, bound, (GATES * H, H))).astype(np.float32)
            sd["gru.bias_ih_l{}{}".format(layer, sfx)] = rs.uniform(
                -bound, bound, (GATES * H,)).astype(np.float32)
            sd["gru.bias_hh_l{}{}".format(layer, sfx)] = rs.uniform(
                -bound, bound, (GATES * H,)).astype(np.float32)
    lb = 1.0 / np.sqrt(H * ndir)
    sd["linear.weight"] = (head_gain * rs.uniform(-lb, lb, (5, H * ndir))).astype(np.float32)
    sd["linear.bias"] = (head_gain * rs.uniform(-lb, lb, (5,))).astype(np.float32)
    return sd


def synth_state_dict_neartie(seed=0, eps=1e-3, bias_shift=-6.318e-4, **kw):
    """Adversarial head for the label-parity tests: classes 1 and 2 get (almost) the same row of the Linear layer and
    (almost) the same, raised, bias - they are the top two everywhere.  ``bias_shift`` centres the logit difference of
    the two on zero (calibrated once for seed 5 with synth_features(6, 400, 10, seed=105): the difference then has
    standard deviation 4e-5), so the top-2 probability margins spread from 0 to ~1e-4: hundreds of positions sit inside
    fp32 re-association noise of flipping, the rest just outside it."""
    sd = synth_state_dict(seed, **kw)
    rs = np.random.RandomState(seed + 7919)
    w = sd["linear.weight"].copy()
    w[2] = w[1] * (1.0 + eps * rs.standard_normal(w.shape[1])).astype(np.float32)
    b = sd["linear.bias"].copy()
    b[1] += 3.0
    b[2] = b[1] + np.float32(bias_shift)
    sd["linear.weight"], sd["linear.bias"] = w.astype(np.float32), b.astype(np.float32)
    return sd


def synth_counts(n_cols, seed=20240923, num_dtypes=1, mean_depth=30, max_depth=120,
                 minor_frac=0.15, start_major=0, start_on_minor=False):
    """Synthetic raw pileup counts the way BASELINE.md section 4 describes them.

    Returns (counts uint64[n_cols, 10*num_dtypes], positions structured
    [('major', i8), ('minor', i8)]) laid out as medaka_counts.c emits them
    (src/medaka_counts.c:274-357): feature order 'acgtACGTdD' per dtype, lower
    case = reverse strand; a major column followed by its minor (insertion)
    columns which only carry the reads that have the insertion.
    """
    rs = np.random.RandomState(seed)
    F = 10 * num_dtypes
    is_minor = rs.uniform(size=n_cols) < minor_frac
    is_minor[0] = bool(start_on_minor)
    major = start_major + np.cumsum(~is_minor) - (0 if start_on_minor else 1)
    major = major.astype(np.int64)
    # minor index = run length since last major
    idx = np.arange(n_cols)
    last_major_idx = np.maximum.accumulate(np.where(~is_minor, idx, -1))
    minor = (idx - last_major_idx).astype(np.int64)
    if start_on_minor:
        # chunk cut in the middle of an insertion run: first column is minor 1..
        minor = np.where(last_major_idx < 0, idx + 1, minor).astype(np.int64)
    depth = np.clip(rs.poisson(mean_depth, n_cols), 1, max_depth)
    counts = np.zeros((n_cols, F), dtype=np.uint64)
    true_base = rs.randint(0, 4, n_cols)
    for dt in range(num_dtypes):
        d_dt = depth if num_dtypes == 1 else rs.binomial(depth, 1.0 / num_dtypes)
        fwd = rs.binomial(d_dt, 0.5)
        rev = d_dt - fwd
        for strand_n, base_off, del_idx in ((rev, 0, 8), (fwd, 4, 9)):
            # 5 outcomes: 4 bases + deletion; 0.94 on the true base
            p = np.full((n_cols, 5), 0.015)
            p[idx, true_base] = 0.94
            p /= p.sum(axis=1, keepdims=True)
            # vectorised multinomial by sequential binomials
            remaining = strand_n.copy()
            rem_p = np.ones(n_cols)
            outs = []
            for k in range(5):
                pk = np.clip(p[:, k] / np.maximum(rem_p, 1e-12), 0, 1)
                x = rs.binomial(remaining, pk) if k < 4 else remaining
                outs.append(x)
                remaining = remaining - x
                rem_p = rem_p - p[:, k]
            for b in range(4):
                counts[:, dt * 10 + base_off + b] = outs[b]
            counts[:, dt * 10 + del_idx] = outs[4]
    # minor columns: only ~15% of the parent's reads carry the insertion, no deletions
    keep = rs.uniform(size=(n_cols, F)) < 0.15
    minor_counts = np.where(keep, counts, 0).astype(np.uint64)
    for dt in range(num_dtypes):
        minor_counts[:, dt * 10 + 8] = 0
        minor_counts[:, dt * 10 + 9] = 0
    counts = np.where(is_minor[:, None] | (minor[:, None] > 0), minor_counts, counts)
    positions = np.empty(n_cols, dtype=[("major", "<i8"), ("minor", "<i8")])
    positions["major"] = major
    positions["minor"] = minor
    return counts, positions


def synth_features(B, T, F=10, seed=1234):
    """Normalised-count-like float32 features [B, T, F] in [0, 1] (sum per dtype <= 1)."""
    rs = np.random.RandomState(seed)
    x = rs.dirichlet(np.full(F, 0.35), size=(B, T)).astype(np.float32)
    # some all-zero columns and sparse insertion-like columns, as real pileups have
    mask = rs.uniform(size=(B, T, 1)) < 0.02
    x = np.where(mask, 0.0, x).astype(np.float32)
    return x


def synth_features_fast(B, T, F=10, seed=1234):
    """Cheap stand-in for synth_features at benchmark sizes (1e8 values in seconds, not minutes):
    cubed uniforms normalised per column, float32 [B, T, F] in [0, 1] with unit row sums."""
    rs = np.random.RandomState(seed)
    x = rs.random_sample((B, T, F)).astype(np.float32)
    x *= x * x
    x /= x.sum(axis=-1, keepdims=True)
    return x


def synth_reads(n_reads, ref_len, seed=0, mean_len=3000, p_ins=0.04, p_del=0.04, p_skip=0.0, num_dtypes=1,
                ins_after_skip=False):
    """Random alignment records (dicts as oracle/pileup_oracle.py takes them): CIGARs with M/=/X runs, insertions
    (incl. consecutive I ops and I right after D), deletions, optional N skips (``ins_after_skip``: insertions may
    follow a skip directly - they widen the column group but are not counted, medaka_counts.c:259-263,282), soft clips,
    both strands, a few filtered flags / low mapQ reads and IUPAC ambiguity codes."""
    rs = np.random.RandomState(seed)
    recs = []
    for i in range(n_reads):
        pos = int(rs.randint(0, max(1, ref_len - 50)))
        target = int(min(ref_len - pos, max(20, rs.exponential(mean_len))))
        ops, ref_used, qlen = [], 0, 0
        if rs.uniform() < 0.3:
            s = int(rs.randint(1, 30)); ops.append((s, "S")); qlen += s
        last = None
        while ref_used < target:
            u = rs.uniform()
            if u < p_ins and last in (("M", "D", "I", "N") if ins_after_skip else ("M", "D", "I")):
                l = int(rs.randint(1, 6)); ops.append((l, "I")); qlen += l; last = "I"
            elif u < p_ins + p_del and last == "M":
                l = int(min(rs.randint(1, 8), target - ref_used)); ops.append((l, "D")); ref_used += l; last = "D"
            elif u < p_ins + p_del + p_skip and last == "M":
                l = int(min(rs.randint(5, 40), target - ref_used)); ops.append((l, "N")); ref_used += l; last = "N"
            else:
                l = int(min(rs.randint(1, 40), target - ref_used))
                ops.append((l, "M=X"[int(rs.randint(0, 3))])); ref_used += l; qlen += l; last = "M"
        if ops[-1][1] not in "M=X":
            ops.append((1, "M")); qlen += 1
        if rs.uniform() < 0.3:
            s = int(rs.randint(1, 30)); ops.append((s, "S")); qlen += s
        alphabet = "ACGT" * 12 + "NRY"
        seq = "".join(alphabet[int(k)] for k in rs.randint(0, len(alphabet), qlen))
        flag = 16 if rs.uniform() < 0.5 else 0
        if rs.uniform() < 0.05:
            flag |= int(rs.choice([0x100, 0x800, 0x400, 0x200]))
        mapq = 0 if rs.uniform() < 0.05 else int(rs.randint(1, 61))
        tags = {"DT": "dt%d" % int(rs.randint(0, num_dtypes))} if num_dtypes > 1 else {}
        recs.append(dict(query_name="r%d" % i, pos=pos, cigar="".join("%d%s" % o for o in ops), seq=seq,
                         flag=flag, mapq=mapq, tags=tags))
    recs.sort(key=lambda r: r["pos"])
    return recs


def synth_stitch_stream(seed=0, n_major=3000, first_major=1000, chunk_len=400, overlap=100, p_ins=0.08,
                        ragged=(), drop=(), nest=(), low_depth=(), ref_name='contig1'):
    """A stream of network-output samples as `medaka consensus` stores them, for the stitching tests.

    One pileup (majors first_major .. first_major+n_major-1, random insertion columns) is cut into overlapping
    chunks like Sample.chunks.  Perturbations, all by chunk index:
      ragged    - delete one insertion column inside the chunk's leading overlap (the two neighbours then disagree
                  on the columns of their overlap -> junction heuristic)
      drop      - remove the chunk from the stream (-> gap between its neighbours when overlap < chunk_len / 2)
      nest      - insert after the chunk a short extra sample lying completely inside it
      low_depth - (chunk, a, b): set depth to 1 on columns [a, b) of the chunk

    :returns: list of dicts {ref_name, positions, label_probs float32 [n,5], depth int64 [n]}.
    """
    rs = np.random.RandomState(seed)
    n_ins = rs.geometric(1.0 - p_ins, size=n_major) - 1
    n_ins = np.minimum(n_ins, 4)
    counts = 1 + n_ins
    major = np.repeat(np.arange(first_major, first_major + n_major), counts)
    starts = np.cumsum(counts) - counts
    minor = np.arange(len(major)) - np.repeat(starts, counts)
    positions = np.empty(len(major), dtype=[('major', int), ('minor', int)])
    positions['major'], positions['minor'] = major, minor
    n = len(positions)

    def make(lo, hi, sub_seed):
        r = np.random.RandomState(sub_seed)
        pos = positions[lo:hi].copy()
        logits = r.randn(hi - lo, 5).astype(np.float32) * 3.0
        logits[pos['minor'] > 0, 0] += 4.0            # insertion columns are mostly gap calls
        logits[pos['minor'] == 0, 0] -= 2.0
        e = np.exp(logits - logits.max(-1, keepdims=True))
        probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
        sure = r.rand(hi - lo) < 0.3                   # confident calls: exercise the quality cap
        probs[sure] = np.eye(5, dtype=np.float32)[np.argmax(probs[sure], -1)] * np.float32(0.99999994) \
            if sure.any() else probs[sure]
        depth = r.randint(20, 60, size=hi - lo).astype(np.int64)
        return dict(ref_name=ref_name, positions=pos, label_probs=probs, depth=depth)

    ranges = []
    step = chunk_len - overlap
    last_end = 0
    for lo in range(0, n - chunk_len + 1, step):
        ranges.append((lo, lo + chunk_len))
        last_end = lo + chunk_len
    if n > last_end:
        ranges.append((max(0, n - chunk_len), n))
    stream = []
    low = {c: (a, b) for c, a, b in low_depth}
    for c, (lo, hi) in enumerate(ranges):
        if c in drop:
            continue
        s = make(lo, hi, seed * 1000 + c)
        if c in ragged:
            cand = np.flatnonzero(s['positions']['minor'][:overlap] > 0)
            # only a trailing insertion column can go without leaving a hole in the minor numbering
            nxt = np.append(s['positions']['minor'][1:], 0)
            cand = [k for k in cand if nxt[k] == 0 and 5 < k < overlap - 5]
            if cand:
                k = cand[len(cand) // 2]
                s = {key: (np.delete(v, k, axis=0) if key != 'ref_name' else v) for key, v in s.items()}
        if c in low:
            a, b = low[c]
            s['depth'][a:b] = 1
        stream.append(s)
        if c in nest:
            stream.append(make(lo + chunk_len // 4, lo + chunk_len // 2, seed * 1000 + 500 + c))
    return stream


def synth_variant_pileup(seed=0, n_major=3000, first_major=1000, p_ins_col=0.08, p_mut=0.03, p_extend=0.35,
                         p_ins_call=0.2, n_frac=0.003, ref_name='contig1'):
    """A pileup with a known draft and a network output that disagrees with it here and there (variant decoding).

    :returns: dict(ref_name, ref_seq (str over 0 .. first_major + n_major + 50, ACGT with a few N), positions
        (major/minor), label_probs float32 [n, 5] whose argmax is the constructed call).

    Calls on major columns: the draft base, or with probability p_mut the start of a mutated run (each further column
    joins with p_extend) of substitutions / deletions; minor (insertion) columns call a base with probability p_ins_call
    (more often right after a mutated major: deletion followed by insertion, the reference's 'CA*cG' case) else gap.
    """
    rs = np.random.RandomState(seed)
    total = first_major + n_major + 50
    ref = rs.randint(0, 4, total)
    is_n = rs.uniform(size=total) < n_frac
    ref_seq = ''.join('N' if is_n[i] else 'ACGT'[ref[i]] for i in range(total))
    n_ins = np.minimum(rs.geometric(1.0 - p_ins_col, size=n_major) - 1, 3)
    counts = 1 + n_ins
    major = np.repeat(np.arange(first_major, first_major + n_major), counts)
    starts = np.cumsum(counts) - counts
    minor = np.arange(len(major)) - np.repeat(starts, counts)
    n = len(major)
    positions = np.empty(n, dtype=[('major', int), ('minor', int)])
    positions['major'], positions['minor'] = major, minor
    call = np.zeros(n, dtype=np.int64)               # label codes: 0 '*', 1..4 ACGT
    in_run = False
    for i in range(n):
        if minor[i] == 0:
            base = ref[major[i]] + 1
            in_run = (rs.uniform() < p_extend) if in_run else (rs.uniform() < p_mut)
            if in_run:
                kind = rs.uniform()
                call[i] = 0 if kind < 0.35 else 1 + (base - 1 + rs.randint(1, 4)) % 4
            else:
                call[i] = base
        else:
            p = 0.6 if in_run else p_ins_call
            call[i] = rs.randint(1, 5) if rs.uniform() < p else 0
    top = rs.uniform(0.45, 0.9999, n)
    top[rs.uniform(size=n) < 0.2] = 0.99999994       # quality cap
    rest = rs.dirichlet(np.ones(4), n) * (1.0 - top)[:, None]
    probs = np.empty((n, 5), dtype=np.float64)
    for c in range(5):
        sel = call == c
        others = [k for k in range(5) if k != c]
        probs[np.ix_(sel, others)] = rest[sel]
        probs[sel, c] = top[sel]
    probs = probs.astype(np.float32)
    # the constructed call must be the unique argmax after the float32 cast
    fix = np.argmax(probs, -1) != call
    probs[fix] = np.eye(5, dtype=np.float32)[call[fix]] * np.float32(0.9)
    probs[fix] += np.float32(0.025)
    assert np.array_equal(np.argmax(probs, -1), call)
    return dict(ref_name=ref_name, ref_seq=ref_seq, positions=positions, label_probs=probs, call=call)
