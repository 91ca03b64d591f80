"""Container types and window scheduling of the hot path: Sample, Region, sliding windows.

Host-side mirror of the parts of medaka/common.py the inference loop uses
(``Sample`` :59-642, ``Region`` :646-759, ``sliding_window`` :803-823, ``grouper``
:903-916).  Same names, argument meaning and error behaviour, written against numpy
only (no pysam / intervaltree), so code written for the reference's types keeps working.
"""
import collections
import enum
import functools
import itertools
import logging
import re

import numpy as np

from medaka_b200 import libmedaka as _lm


def get_named_logger(name):
    """medaka/common.py:937-941."""
    name = name.ljust(10)[:10]
    return logging.getLogger("{}.{}".format(__package__, name))


def _plp_bases():
    lib = _lm.load()
    return _lm.ffi.string(lib.mdk_plp_bases()).decode()


@functools.lru_cache(maxsize=1)
def base2index():
    """Mapping from base symbol to column of the counts matrix (medaka/common.py:29-35)."""
    return {c: i for i, c in enumerate(_plp_bases())}


class OverlapException(Exception):
    """Exception class used when examining range overlaps."""


class Relationship(enum.Enum):
    """Enumeration of types of overlap (medaka/common.py:44-55)."""

    different_ref_name = 'Samples come from different reference contigs.'
    forward_overlap = 'The end of s1 overlaps the start of s2.'
    reverse_overlap = 'The end of s2 overlaps the start of s1.'
    forward_abutted = 'The end of s1 abuts the start of s2.'
    reverse_abutted = 'The end of s2 abuts the start of s1.'
    forward_gapped = 's2 follows s1 with a gab inbetween.'
    reverse_gapped = 's1 follows s2 with a gab inbetween.'
    s2_within_s1 = 's2 is fully contained within s1.'
    s1_within_s2 = 's1 is fully contained within s2.'


_Sample = collections.namedtuple(
    'Sample',
    ['ref_name', 'features', 'labels', 'ref_seq', 'positions', 'label_probs', 'depth'])


class Sample(_Sample):
    """A pileup range: the unit the hot path consumes (features) and produces (label_probs)."""

    def _asdict(self):
        return collections.OrderedDict(zip(self._fields, self))

    def amend(self, **kwargs):
        """Create new `Sample` with some attributes changed."""
        d = self._asdict()
        for k, v in kwargs.items():
            if k not in self._fields:
                raise KeyError('Invalid key for Sample: {}'.format(k))
            d[k] = v
        return Sample(**d)

    def _get_pos(self, index):
        p = self.positions
        return p['major'][index], p['minor'][index]

    @property
    def first_pos(self):
        """Zero-based first reference co-ordinate."""
        return self._get_pos(0)

    @property
    def last_pos(self):
        """Zero-based (end inclusive) last reference co-ordinate."""
        return self._get_pos(-1)

    @property
    def span(self):
        """Size of sample in terms of reference positions."""
        return self.last_pos[0] - self.first_pos[0]

    @property
    def is_empty(self):
        """Is pileup empty, synonymous to `sample.size == 0`."""
        return self.size == 0

    @property
    def size(self):
        """Return number of columns of pileup."""
        return len(self.positions)

    @property
    def name(self):
        """Create zero-based (end inclusive) samtools-style region string."""
        fmaj, fmin = self.first_pos
        lmaj, lmin = self.last_pos
        return '{}:{}.{}-{}.{}'.format(self.ref_name, fmaj, fmin, lmaj, lmin)

    @staticmethod
    def decode_sample_name(name):
        """Decode the result of Sample.name into a dict."""
        d = None
        m = re.match(r"(?P<ref_name>.+):(?P<start>\d+\.\d+)-(?P<end>\d+\.\d+)", name)
        if m is not None:
            d = m.groupdict()
        return d

    @staticmethod
    def from_samples(samples):
        """Concatenate abutting samples into one (medaka/common.py:200-230)."""
        samples = list(samples)
        for s1, s2 in zip(samples[:-1], samples[1:]):
            rel = Sample.relative_position(s1, s2)
            if rel is not Relationship.forward_abutted:
                raise ValueError('Refusing to concatenate unordered/non-abutting samples {} and {} with relationship {}.'
                                 .format(s1.name, s2.name, repr(rel)))
        fields = {}
        for attr in Sample._fields:
            vals = [getattr(s, attr) for s in samples]
            if attr == 'ref_name':
                fields[attr] = vals[0]
            else:
                fields[attr] = None if all(v is None for v in vals) else np.concatenate(vals)
        return Sample(**fields)

    @staticmethod
    def relative_position(s1, s2):
        """Classify how two samples sit relative to each other (cf. medaka/common.py:231-325).

        Positions are compared as (major, minor) pairs; the pair that starts first (ties: the
        longer one) is taken as the left sample.
        """
        if s1.ref_name != s2.ref_name:
            return Relationship.different_ref_name

        def key(s):
            return (tuple(int(v) for v in s.first_pos), -s.size)

        swapped = key(s2) < key(s1)
        left, right = (s2, s1) if swapped else (s1, s2)
        l_first, l_last = tuple(map(int, left.first_pos)), tuple(map(int, left.last_pos))
        r_first, r_last = tuple(map(int, right.first_pos)), tuple(map(int, right.last_pos))
        if r_first >= l_first and r_last <= l_last:
            kind = 'within'
        elif r_first in ((l_last[0] + 1, 0), (l_last[0], l_last[1] + 1)):
            kind = 'abutted'
        elif r_first <= l_last:
            kind = 'overlap'
        else:
            kind = 'gapped'
        table = {
            ('within', False): Relationship.s2_within_s1, ('within', True): Relationship.s1_within_s2,
            ('abutted', False): Relationship.forward_abutted, ('abutted', True): Relationship.reverse_abutted,
            ('overlap', False): Relationship.forward_overlap, ('overlap', True): Relationship.reverse_overlap,
            ('gapped', False): Relationship.forward_gapped, ('gapped', True): Relationship.reverse_gapped,
        }
        return table[kind, swapped]

    def chunks(self, chunk_len=1000, overlap=200):
        """Overlapping windows of self, as views (semantics of medaka/common.py:429-453).

        Windows start every ``chunk_len - overlap`` columns; a trailing remainder is covered by one
        extra window right-aligned to the end, so every yielded window has exactly ``chunk_len``
        columns.  A sample shorter than ``chunk_len`` yields itself once (callers quarantine those).
        """
        for start, end in window_ranges(self.size, chunk_len, chunk_len - overlap):
            yield self.slice(slice(start, end))

    def slice(self, key):
        """Slice fields along the genomic axis (views of the original arrays)."""
        def slice_attr(attr):
            a = getattr(self, attr)
            if attr != 'ref_name':
                a = a[key] if a is not None else None
            return a
        return Sample(**{attr: slice_attr(attr) for attr in self._fields})

    def __eq__(self, other):
        """Test equality."""
        for field in self._fields:
            s = getattr(self, field)
            o = getattr(other, field)
            if type(s) is not type(o):
                return False
            elif isinstance(s, np.ndarray):
                if (s.shape != o.shape or np.any(s != o)):
                    return False
            elif s != o:
                return False
        return True

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None


_Region = collections.namedtuple('Region', 'ref_name start end')


class Region(_Region):
    """Represents a genomic region (medaka/common.py:649-759)."""

    @property
    def name(self):
        """Samtools-style region string, zero-base end exclusive."""
        return self.__str__()

    def __str__(self):
        start = 0 if self.start is None else self.start
        end = '' if self.end is None else self.end
        return '{}:{}-{}'.format(self.ref_name, start, end)

    @property
    def size(self):
        """Return size of region."""
        return self.end - self.start

    @classmethod
    def from_string(cls, region):
        """Parse region string into `Region` objects.

        >>> Region.from_string('Ecoli:1000-2000') == Region('Ecoli', 1000, 2000)
        True
        >>> Region.from_string('A:B:c:500-') == Region('A:B:c', 500, None)
        True
        """
        if ':' not in region:
            ref_name, start, end = region, None, None
        else:
            start, end = None, None
            ref_name, bounds = region.rsplit(':', 1)
            if bounds[0] == '-':
                start = 0
                end = int(bounds.replace('-', ''))
            elif '-' not in bounds:
                start = int(bounds)
                end = None
            elif bounds[-1] == '-':
                start = int(bounds[:-1])
                end = None
            else:
                start, end = [int(b) for b in bounds.split('-')]
        return cls(ref_name, start, end)

    def split(region, size, overlap=0, fixed_size=True):
        """Split region into sub-regions of a given length (medaka/common.py:712-737)."""
        regions = list()
        if size >= region.size:
            return [region]
        for start in range(region.start, region.end, size - overlap):
            end = min(start + size, region.end)
            regions.append(Region(region.ref_name, start, end))
        if len(regions) > 1:
            if fixed_size and regions[-1].size < size:
                del regions[-1]
                end = region.end
                start = end - size
                if start > regions[-1].start:
                    regions.append(Region(region.ref_name, start, end))
        return regions

    def overlaps(self, other):
        """Determine if a region overlaps another."""
        if self.ref_name != other.ref_name:
            return False

        def _limits(x):
            x0 = x.start if x.start is not None else -1
            x1 = x.end if x.end is not None else float('inf')
            return x0, x1

        a0, a1 = _limits(self)
        b0, b1 = _limits(other)
        return (a0 < b1 and a1 > b0) or (b0 < a1 and b1 > a0)


def window_ranges(n, window, step):
    """(start, end) index pairs of the windows ``sliding_window`` visits over an axis of length n."""
    starts = list(range(0, max(n - window + 1, 0), step))
    ranges = [(s0, s0 + window) for s0 in starts]
    covered = ranges[-1][1] if ranges else 0
    if n > covered:
        # right-aligned remainder window; for n < window numpy-style slicing clamps it to [0, n)
        ranges.append((max(n - window, 0) if n >= window else n - window, n))
    return ranges


def sliding_window(a, window=3, step=1, axis=0):
    """Windows of ``a`` along ``axis``; the remainder is yielded right-aligned (cf. common.py:803-823)."""
    n = a.shape[axis]
    index = [slice(None)] * a.ndim
    for start, end in window_ranges(n, window, step):
        index[axis] = slice(start, end)
        yield a[tuple(index)]


def grouper(gen, batch_size=4):
    """Group together elements of an iterable without padding remainder (common.py:903-916)."""
    if not isinstance(gen, collections.abc.Iterator):
        gen = iter(gen)
    while True:
        batch = []
        for i in range(batch_size):
            try:
                batch.append(next(gen))
            except StopIteration:
                if len(batch) > 0:
                    yield batch
                return
        yield batch


def rle(array, low_mem=False):
    """Run-length encode a 1-D array (medaka/common.py:1125-1147) -> structured (length, start, value)."""
    if len(array.shape) != 1:
        raise TypeError("Input array must be one dimensional.")
    dtype = [('length', int), ('start', int), ('value', array.dtype)]
    n = len(array)
    if n == 0:
        return np.empty(0, dtype=dtype)
    starts = np.r_[0, np.flatnonzero(array[1:] != array[:-1]) + 1]
    lengths = np.diff(np.r_[starts, n])
    return np.fromiter(zip(lengths, starts, array[starts]), dtype, len(lengths))
