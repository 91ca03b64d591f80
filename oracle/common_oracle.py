"""Oracle for window scheduling: sliding_window / Sample.chunks / Region.split / grouper.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows medaka/common.py:803-823 (sliding_window: stride = window-overlap, last
window right-aligned), :429-453 (Sample.chunks), :712-737 (Region.split) and
:903-916 (grouper).  Expressed on index ranges so it can be compared with the host
scheduler in medaka_b200.common without building Sample objects.
"""


def sliding_window_ranges(n, window, step):
    """(start, end) of every window medaka.common.sliding_window yields over axis length n."""
    out = []
    end = 0
    for start in range(0, n - window + 1, step):
        end = start + window
        out.append((start, end))
    if n > end:
        out.append((n - window, n))
    return out


def chunk_ranges(n, chunk_len, overlap):
    """Index ranges of Sample.chunks(chunk_len, overlap) for a sample of n columns."""
    return sliding_window_ranges(n, chunk_len, chunk_len - overlap)


def region_split(start, end, size, overlap=0, fixed_size=True):
    """medaka/common.py:712-737 on (start, end) tuples."""
    if size >= end - start:
        return [(start, end)]
    regions = []
    for s in range(start, end, size - overlap):
        regions.append((s, min(s + size, end)))
    if len(regions) > 1:
        if fixed_size and regions[-1][1] - regions[-1][0] < size:
            del regions[-1]
            s = end - size
            if s > regions[-1][0]:
                regions.append((s, end))
    return regions


def grouper(items, batch_size):
    items = list(items)
    return [items[i:i + batch_size] for i in range(0, len(items), batch_size)]
