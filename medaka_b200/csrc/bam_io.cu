// medaka_b200: native BAM access for the GPU pileup featuriser (host code only; compiled into libmedaka_b200.so).
//
// The reference reads alignments through htslib: create_bam_fset opens the file and its index
// (src/medaka_bamiter.c:52-63), calculate_pileup asks for an iterator over the region (bam_itr_querys,
// src/medaka_counts.c:233) and read_bam (src/medaka_bamiter.c:8-48) filters the records that iterator yields.  htslib
// is not part of the reference tree (downloaded at build time) and not in this image, so this file implements what
// that path needs from the BAM / BGZF / BAI formats (SAM specification, sections 4.1, 4.2, 5.2):
//   * BGZF: the file is a series of <= 64 KiB gzip members; a virtual offset is (member file offset << 16 | offset
//     inside the inflated member).  Members inflate independently -> a thread pool (zlib raw inflate).
//   * BAI: per reference, bins -> chunks of virtual offsets plus a 16 kb linear index; reg2bins + the linear-index
//     lower bound give the byte ranges that can hold records overlapping [start, end).
//   * without a .bai the file is streamed once, member by member, in bounded memory.
// fetch() returns the records overlapping the region that pass the flag / mapping-quality part of read_bam
// (medaka_bamiter.c:19-21 - these run BEFORE the tag filters there, so they do here), in BAM's own packed encodings
// (32-bit CIGAR ops, 4-bit bases) as flat arrays ready for mdk_pileup_counts, plus each record's aux bytes for the
// tag / read-group / datatype filters of the caller (medaka_b200/bam.py).
// Long CIGARs: a read with more than 65535 operations stores the placeholder <l_seq>S<ref_len>N and its real CIGAR in
// the CG:B,I tag (SAM spec 4.2.2); htslib swaps it in transparently when it reads the record, so does fetch().
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

struct mdk_bam {
    std::string path;
    FILE *fp = nullptr;
    int64_t file_size = 0;
    std::vector<std::string> ref_names;
    std::vector<int32_t> ref_lens;
    uint64_t first_voffset = 0;           // virtual offset of the first alignment record
    bool has_index = false;
    struct RefIndex {
        std::unordered_map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
        std::vector<uint64_t> linear;
    };
    std::vector<RefIndex> index;
    std::mutex io_mutex;                  // fetch() may be called from several Python threads on one handle
};

struct mdk_bam_batch {
    std::vector<int32_t> pos, l_seq;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq;
    std::vector<uint32_t> cigar;
    std::vector<int64_t> cigar_off{0};
    std::vector<uint8_t> seq;
    std::vector<int64_t> seq_off{0};
    std::vector<uint8_t> qual;          // l_seq bytes per read (0xff = absent), offsets qual_off
    std::vector<int64_t> qual_off{0};
    std::vector<uint8_t> aux;
    std::vector<int64_t> aux_off{0};
    std::vector<char> names;
    std::vector<int64_t> name_off{0};
};

namespace {

using mdk::set_error;

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// one BGZF member inside a buffer of compressed bytes
struct Member {
    size_t c_off;      // offset of the member in the compressed buffer
    uint32_t bsize;    // total member size
    uint32_t xlen;
    uint32_t isize;    // inflated size
    size_t u_off;      // offset of its data in the inflated buffer
};

// Walk the members of buf[0, n).  Returns false on a malformed header; a trailing partial member is left out (*used
// reports how many bytes were consumed).
bool scan_members(const uint8_t *buf, size_t n, std::vector<Member> &out, size_t *used) {
    size_t off = 0, u = 0;
    while (off + 18 <= n) {
        if (!(buf[off] == 0x1f && buf[off + 1] == 0x8b && buf[off + 2] == 8 && (buf[off + 3] & 4))) return false;
        const uint32_t xlen = rd16(buf + off + 10);
        if (off + 12 + xlen > n) break;
        uint32_t bsize = 0;
        bool found = false;
        for (size_t x = off + 12; x + 4 <= off + 12 + xlen;) {
            const uint32_t slen = rd16(buf + x + 2);
            if (buf[x] == 66 && buf[x + 1] == 67 && slen == 2) { bsize = (uint32_t)rd16(buf + x + 4) + 1; found = true; }
            x += 4 + slen;
        }
        if (!found) return false;
        if (off + bsize > n) break;
        const uint32_t isize = rd32(buf + off + bsize - 4);
        out.push_back(Member{off, bsize, xlen, isize, u});
        u += isize;
        off += bsize;
    }
    *used = off;
    return true;
}

bool inflate_member(const uint8_t *buf, const Member &m, uint8_t *dst) {
    if (m.isize == 0) return true;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<uint8_t *>(buf + m.c_off + 12 + m.xlen);
    zs.avail_in = m.bsize - m.xlen - 20;
    zs.next_out = dst;
    zs.avail_out = m.isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    return rc == Z_STREAM_END && zs.avail_out == 0;
}

bool inflate_all(const uint8_t *buf, const std::vector<Member> &ms, std::vector<uint8_t> &out, int threads) {
    size_t total = 0;
    for (const Member &m : ms) total += m.isize;
    out.resize(total);
    if (threads < 1) threads = 1;
    if (threads > 1 && ms.size() >= 8) {
        std::vector<std::thread> pool;
        std::vector<char> ok((size_t)threads, 1);
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([&, t]() {
                for (size_t i = (size_t)t; i < ms.size(); i += (size_t)threads)
                    if (!inflate_member(buf, ms[i], out.data() + ms[i].u_off)) ok[(size_t)t] = 0;
            });
        for (auto &th : pool) th.join();
        for (char c : ok) if (!c) return false;
        return true;
    }
    for (const Member &m : ms)
        if (!inflate_member(buf, m, out.data() + m.u_off)) return false;
    return true;
}

bool read_at(mdk_bam *b, int64_t off, size_t n, std::vector<uint8_t> &dst) {
    dst.resize(n);
    if (fseeko(b->fp, (off_t)off, SEEK_SET) != 0) return false;
    const size_t got = fread(dst.data(), 1, n, b->fp);
    dst.resize(got);
    return true;
}

// reg2bins of the SAM specification (5.3): the bins that may hold records overlapping [beg, end)
void reg2bins(int64_t beg, int64_t end, std::vector<uint32_t> &bins) {
    --end;
    bins.push_back(0);
    for (uint32_t k = 1 + (uint32_t)(beg >> 26); k <= 1 + (uint32_t)(end >> 26); ++k) bins.push_back(k);
    for (uint32_t k = 9 + (uint32_t)(beg >> 23); k <= 9 + (uint32_t)(end >> 23); ++k) bins.push_back(k);
    for (uint32_t k = 73 + (uint32_t)(beg >> 20); k <= 73 + (uint32_t)(end >> 20); ++k) bins.push_back(k);
    for (uint32_t k = 585 + (uint32_t)(beg >> 17); k <= 585 + (uint32_t)(end >> 17); ++k) bins.push_back(k);
    for (uint32_t k = 4681 + (uint32_t)(beg >> 14); k <= 4681 + (uint32_t)(end >> 14); ++k) bins.push_back(k);
}

bool load_bai(mdk_bam *b, const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::vector<uint8_t> d;
    fseeko(f, 0, SEEK_END);
    const off_t sz = ftello(f);
    fseeko(f, 0, SEEK_SET);
    d.resize((size_t)sz);
    const size_t got = fread(d.data(), 1, d.size(), f);
    fclose(f);
    if (got != d.size() || d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return false;
    size_t o = 4;
    const uint32_t n_ref = rd32(d.data() + o);
    o += 4;
    b->index.assign(n_ref, mdk_bam::RefIndex());
    for (uint32_t r = 0; r < n_ref; ++r) {
        if (o + 4 > d.size()) return false;
        const uint32_t n_bin = rd32(d.data() + o);
        o += 4;
        for (uint32_t k = 0; k < n_bin; ++k) {
            if (o + 8 > d.size()) return false;
            const uint32_t bin = rd32(d.data() + o), n_chunk = rd32(d.data() + o + 4);
            o += 8;
            if (o + (size_t)n_chunk * 16 > d.size()) return false;
            if (bin != 37450) {           // 37450 is the metadata pseudo-bin
                auto &v = b->index[r].bins[bin];
                for (uint32_t c = 0; c < n_chunk; ++c) v.emplace_back(rd64(d.data() + o + c * 16), rd64(d.data() + o + c * 16 + 8));
            }
            o += (size_t)n_chunk * 16;
        }
        if (o + 4 > d.size()) return false;
        const uint32_t n_intv = rd32(d.data() + o);
        o += 4;
        if (o + (size_t)n_intv * 8 > d.size()) return false;
        b->index[r].linear.resize(n_intv);
        for (uint32_t i = 0; i < n_intv; ++i) b->index[r].linear[i] = rd64(d.data() + o + i * 8);
        o += (size_t)n_intv * 8;
    }
    return true;
}

constexpr int CONSUMES_REF[16] = {1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0};   // M I D N S H P = X

// Look for the CG:B,I tag in a record's aux bytes; returns a pointer to its uint32 array and the count
const uint8_t *find_cg(const uint8_t *aux, size_t n, uint32_t *count) {
    size_t i = 0;
    while (i + 3 <= n) {
        const uint8_t t0 = aux[i], t1 = aux[i + 1], typ = aux[i + 2];
        i += 3;
        size_t sz = 0;
        switch (typ) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': {
                const void *z = memchr(aux + i, 0, n - i);
                if (!z) return nullptr;
                sz = (size_t)((const uint8_t *)z - (aux + i)) + 1;
                break;
            }
            case 'B': {
                if (i + 5 > n) return nullptr;
                const uint8_t sub = aux[i];
                const uint32_t cnt = rd32(aux + i + 1);
                const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                if (t0 == 'C' && t1 == 'G' && (sub == 'I' || sub == 'i')) {
                    if (i + 5 + (size_t)cnt * 4 > n) return nullptr;
                    *count = cnt;
                    return aux + i + 5;
                }
                sz = 5 + (size_t)cnt * w;
                break;
            }
            default: return nullptr;
        }
        i += sz;
    }
    return nullptr;
}

// Append the records of data[u0, u1) that are on `tid`, overlap [start, end) and pass the flag / mapq filters.
// Returns the offset where parsing stopped (start of the first incomplete record) or -1 on a malformed record.
// *past_end is set when a record beyond the region (pos >= end, or a later reference) was seen: BAM files the pileup
// reads are coordinate sorted, so nothing after it can overlap.
int64_t parse_records(const uint8_t *data, int64_t u0, int64_t u1, int tid, int32_t start, int32_t end,
                      uint32_t exclude_flags, int min_mapq, mdk_bam_batch *out, bool *past_end) {
    int64_t off = u0;
    while (off + 4 <= u1) {
        const int32_t bs = (int32_t)rd32(data + off);
        if (bs < 32) return -1;
        if (off + 4 + bs > u1) break;
        const uint8_t *r = data + off + 4;
        const int32_t ref_id = (int32_t)rd32(r), pos = (int32_t)rd32(r + 4);
        const uint32_t l_read_name = r[8], mapq = r[9], n_cigar = rd16(r + 12), flag = rd16(r + 14);
        const int32_t l_seq = (int32_t)rd32(r + 16);
        const int64_t fixed = 32 + (int64_t)l_read_name + 4 * (int64_t)n_cigar + (l_seq + 1) / 2 + l_seq;
        if (l_seq < 0 || fixed > bs) return -1;
        off += 4 + bs;
        if (ref_id != tid) {
            if (ref_id > tid || ref_id < 0) { *past_end = true; return off; }
            continue;
        }
        if (pos >= end) { *past_end = true; return off; }
        const uint8_t *cig = r + 32 + l_read_name;
        const uint8_t *seq = cig + 4 * n_cigar;
        const uint8_t *aux = seq + (l_seq + 1) / 2 + l_seq;
        const size_t aux_len = (size_t)(bs - fixed);
        const uint8_t *ops = cig;
        uint32_t n_ops = n_cigar;
        if (n_cigar == 2) {   // long-CIGAR placeholder <l_seq>S<ref_len>N: the real operations are in CG:B,I
            const uint32_t op0 = rd32(cig), op1 = rd32(cig + 4);
            if ((op0 & 15) == 4 && (int32_t)(op0 >> 4) == l_seq && (op1 & 15) == 3) {
                uint32_t cnt = 0;
                const uint8_t *cg = find_cg(aux, aux_len, &cnt);
                if (cg && cnt > 0) { ops = cg; n_ops = cnt; }
            }
        }
        int64_t ref_len = 0;
        for (uint32_t k = 0; k < n_ops; ++k) {
            const uint32_t op = rd32(ops + 4 * k);
            ref_len += (int64_t)(op >> 4) * CONSUMES_REF[op & 15];
        }
        if ((int64_t)pos + ref_len <= start) continue;           // ends before the region
        if (flag & exclude_flags) continue;                      // medaka_bamiter.c:19
        if ((int)mapq < min_mapq) continue;                      // medaka_bamiter.c:21
        out->pos.push_back(pos);
        out->flag.push_back((uint16_t)flag);
        out->mapq.push_back((uint8_t)mapq);
        out->l_seq.push_back(l_seq);
        for (uint32_t k = 0; k < n_ops; ++k) out->cigar.push_back(rd32(ops + 4 * k));
        out->cigar_off.push_back((int64_t)out->cigar.size());
        out->seq.insert(out->seq.end(), seq, seq + (l_seq + 1) / 2);
        out->seq_off.push_back((int64_t)out->seq.size());
        out->qual.insert(out->qual.end(), seq + (l_seq + 1) / 2, seq + (l_seq + 1) / 2 + l_seq);
        out->qual_off.push_back((int64_t)out->qual.size());
        out->aux.insert(out->aux.end(), aux, aux + aux_len);
        out->aux_off.push_back((int64_t)out->aux.size());
        const char *nm = reinterpret_cast<const char *>(r + 32);
        out->names.insert(out->names.end(), nm, nm + (l_read_name ? l_read_name - 1 : 0));
        out->name_off.push_back((int64_t)out->names.size());
    }
    return off;
}

// Inflate the compressed range [c0, c1) of the file (member aligned at c0; c1 is clipped to the file) in pieces of
// at most `piece` compressed bytes and feed the inflated stream to parse_records, carrying a record that straddles
// two pieces.  first_uoff = offset of the first wanted byte inside the first member.
int stream_range(mdk_bam *b, int64_t c0, int64_t c1, uint32_t first_uoff, int tid, int32_t start, int32_t end,
                 uint32_t exclude_flags, int min_mapq, int threads, mdk_bam_batch *out, bool stop_at_region_end) {
    const size_t piece = (size_t)8 << 20;
    std::vector<uint8_t> cbuf, ubuf, carry;
    int64_t off = c0;
    bool first = true, past_end = false;
    if (c1 > b->file_size) c1 = b->file_size;
    while (off < c1 && !(past_end && stop_at_region_end)) {
        // a record may straddle the end of the wanted range: always read one more member's worth beyond c1
        const size_t want = (size_t)std::min<int64_t>((int64_t)piece, c1 - off) + 65536 + 4096;
        {
            std::lock_guard<std::mutex> lock(b->io_mutex);
            if (!read_at(b, off, want, cbuf)) { set_error("bam: read failed: " + b->path); return MDK_ERR_ARG; }
        }
        if (cbuf.empty()) break;
        std::vector<Member> ms;
        size_t used = 0;
        if (!scan_members(cbuf.data(), cbuf.size(), ms, &used) || ms.empty()) {
            set_error("bam: malformed BGZF member in " + b->path);
            return MDK_ERR_ARG;
        }
        // keep only whole members that start before c1 (the extra bytes were read for the last of them)
        while (ms.size() > 1 && (int64_t)(off + (int64_t)ms.back().c_off) >= c1) ms.pop_back();
        if (!inflate_all(cbuf.data(), ms, ubuf, threads)) { set_error("bam: inflate failed in " + b->path); return MDK_ERR_ARG; }
        const size_t skip = first ? first_uoff : 0;
        first = false;
        if (skip > ubuf.size()) { set_error("bam: virtual offset beyond its BGZF member"); return MDK_ERR_ARG; }
        carry.insert(carry.end(), ubuf.begin() + (std::ptrdiff_t)skip, ubuf.end());
        const int64_t stopped = parse_records(carry.data(), 0, (int64_t)carry.size(), tid, start, end, exclude_flags,
                                              min_mapq, out, &past_end);
        if (stopped < 0) { set_error("bam: malformed alignment record in " + b->path); return MDK_ERR_ARG; }
        carry.erase(carry.begin(), carry.begin() + (std::ptrdiff_t)stopped);
        off += (int64_t)(ms.back().c_off + ms.back().bsize);
    }
    return MDK_OK;
}

}  // namespace

extern "C" {

int mdk_bam_open(const char *path, const char *index_path, mdk_bam **out) {
    MDK_REQUIRE(path && out, MDK_ERR_ARG, "bam_open: NULL argument");
    mdk_bam *b = new (std::nothrow) mdk_bam();
    MDK_REQUIRE(b, MDK_ERR_NOMEM, "bam_open: out of host memory");
    b->path = path;
    b->fp = fopen(path, "rb");
    if (!b->fp) { delete b; set_error(std::string("bam_open: cannot open ") + path); return MDK_ERR_ARG; }
    fseeko(b->fp, 0, SEEK_END);
    b->file_size = (int64_t)ftello(b->fp);
    // header: inflate members from the start until magic, text, and the reference table are complete
    std::vector<uint8_t> cbuf, ubuf, hdr;
    int64_t coff = 0;
    std::vector<std::pair<int64_t, size_t>> member_at;   // (file offset, inflated bytes before it) of header members
    size_t need = 12;
    bool done = false;
    size_t parsed_refs = 0, o = 0;
    int32_t n_ref = -1;
    while (!done) {
        if (coff >= b->file_size) { mdk_bam_close(b); set_error("bam_open: truncated header"); return MDK_ERR_ARG; }
        if (!read_at(b, coff, (size_t)1 << 20, cbuf)) { mdk_bam_close(b); set_error("bam_open: read failed"); return MDK_ERR_ARG; }
        std::vector<Member> ms;
        size_t used = 0;
        if (!scan_members(cbuf.data(), cbuf.size(), ms, &used) || ms.empty()) {
            mdk_bam_close(b);
            set_error(std::string("bam_open: not a BGZF file: ") + path);
            return MDK_ERR_ARG;
        }
        if (!inflate_all(cbuf.data(), ms, ubuf, 1)) { mdk_bam_close(b); set_error("bam_open: inflate failed"); return MDK_ERR_ARG; }
        for (const Member &m : ms) member_at.emplace_back(coff + (int64_t)m.c_off, hdr.size() + m.u_off);
        hdr.insert(hdr.end(), ubuf.begin(), ubuf.end());
        coff += (int64_t)used;
        // try to parse what we have
        while (true) {
            if (hdr.size() < need) break;
            if (n_ref < 0) {
                if (memcmp(hdr.data(), "BAM\1", 4) != 0) { mdk_bam_close(b); set_error(std::string("bam_open: not a BAM file: ") + path); return MDK_ERR_ARG; }
                const uint32_t l_text = rd32(hdr.data() + 4);
                need = 8 + (size_t)l_text + 4;
                if (hdr.size() < need) break;
                n_ref = (int32_t)rd32(hdr.data() + 8 + l_text);
                o = 8 + (size_t)l_text + 4;
                need = o;
            }
            if ((int32_t)parsed_refs == n_ref) { done = true; break; }
            need = o + 4;
            if (hdr.size() < need) break;
            const uint32_t l_name = rd32(hdr.data() + o);
            need = o + 4 + l_name + 4;
            if (hdr.size() < need) break;
            b->ref_names.emplace_back(reinterpret_cast<const char *>(hdr.data() + o + 4), l_name ? l_name - 1 : 0);
            b->ref_lens.push_back((int32_t)rd32(hdr.data() + o + 4 + l_name));
            o += 8 + l_name;
            need = o;
            ++parsed_refs;
        }
    }
    // virtual offset of the first record: the member that holds inflated byte `o`
    {
        size_t k = 0;
        while (k + 1 < member_at.size() && member_at[k + 1].second <= o) ++k;
        b->first_voffset = ((uint64_t)member_at[k].first << 16) | (uint64_t)(o - member_at[k].second);
        // (a header ending exactly at a member boundary: offset == isize of that member; stream_range copes, the
        // member contributes no bytes)
    }
    std::string ip = index_path ? std::string(index_path) : (b->path + ".bai");
    b->has_index = load_bai(b, ip);
    if (!b->has_index && !index_path) {
        // sample.bam -> sample.bai
        const size_t dot = b->path.rfind(".bam");
        if (dot != std::string::npos && dot + 4 == b->path.size()) b->has_index = load_bai(b, b->path.substr(0, dot) + ".bai");
    }
    if (b->has_index && b->index.size() != b->ref_names.size()) { b->has_index = false; b->index.clear(); }
    *out = b;
    return MDK_OK;
}

int mdk_bam_close(mdk_bam *b) {
    if (!b) return MDK_OK;
    if (b->fp) fclose(b->fp);
    delete b;
    return MDK_OK;
}

int mdk_bam_n_refs(mdk_bam *b) { return b ? (int)b->ref_names.size() : 0; }
const char *mdk_bam_ref_name(mdk_bam *b, int i) {
    return (b && i >= 0 && i < (int)b->ref_names.size()) ? b->ref_names[(size_t)i].c_str() : "";
}
int32_t mdk_bam_ref_len(mdk_bam *b, int i) { return (b && i >= 0 && i < (int)b->ref_lens.size()) ? b->ref_lens[(size_t)i] : 0; }
int mdk_bam_has_index(mdk_bam *b) { return b && b->has_index ? 1 : 0; }

int mdk_bam_fetch(mdk_bam *b, int tid, int32_t start, int32_t end, uint32_t exclude_flags, int min_mapq, int threads,
                  mdk_bam_batch **out) {
    MDK_REQUIRE(b && out, MDK_ERR_ARG, "bam_fetch: NULL argument");
    MDK_REQUIRE(tid >= 0 && tid < (int)b->ref_names.size(), MDK_ERR_ARG, "bam_fetch: no such reference sequence");
    MDK_REQUIRE(start >= 0 && end >= start, MDK_ERR_ARG, "bam_fetch: bad region");
    mdk_bam_batch *batch = new (std::nothrow) mdk_bam_batch();
    MDK_REQUIRE(batch, MDK_ERR_NOMEM, "bam_fetch: out of host memory");
    int rc = MDK_OK;
    if (end > start) {
        if (b->has_index) {
            // candidate chunks: every chunk of every bin overlapping the region whose end lies beyond the linear
            // index' lower bound for `start`; merged into disjoint ranges of virtual offsets
            const mdk_bam::RefIndex &ri = b->index[(size_t)tid];
            uint64_t min_off = 0;
            if (!ri.linear.empty()) {
                const size_t w = std::min<size_t>((size_t)(start >> 14), ri.linear.size() - 1);
                min_off = ri.linear[w];
            }
            std::vector<uint32_t> bins;
            reg2bins(start, end, bins);
            std::vector<std::pair<uint64_t, uint64_t>> chunks;
            for (uint32_t bin : bins) {
                auto it = ri.bins.find(bin);
                if (it == ri.bins.end()) continue;
                for (const auto &c : it->second)
                    if (c.second > min_off) chunks.emplace_back(std::max(c.first, min_off), c.second);
            }
            std::sort(chunks.begin(), chunks.end());
            std::vector<std::pair<uint64_t, uint64_t>> merged;
            for (const auto &c : chunks) {
                // chunks whose compressed ranges touch are read as one (they share BGZF members)
                if (!merged.empty() && (c.first >> 16) <= (merged.back().second >> 16)) merged.back().second = std::max(merged.back().second, c.second);
                else merged.push_back(c);
            }
            for (const auto &c : merged) {
                // read up to and including the member holding the chunk end; records are parsed to the end of that
                // member, the overlap / sorted-order tests discard what lies outside the region
                rc = stream_range(b, (int64_t)(c.first >> 16), (int64_t)(c.second >> 16) + 1, (uint32_t)(c.first & 0xffff),
                                  tid, start, end, exclude_flags, min_mapq, threads, batch, true);
                if (rc != MDK_OK) break;
            }
            // merged ranges can still hand the same record twice when two chunks end / begin inside one member
            if (rc == MDK_OK && merged.size() > 1) {
                // records are identified by (pos, name): drop exact neighbours' duplicates after a stable sort by input order
                // (duplicates are adjacent range boundaries only); cheap O(n) pass on names
                std::vector<size_t> keep;
                const size_t n = batch->pos.size();
                std::unordered_map<std::string, int> seen;
                bool dup = false;
                std::vector<char> is_dup(n, 0);
                for (size_t i = 0; i < n; ++i) {
                    std::string key(batch->names.data() + batch->name_off[i], (size_t)(batch->name_off[i + 1] - batch->name_off[i]));
                    key += ':' + std::to_string(batch->pos[i]) + ':' + std::to_string(batch->flag[i]);
                    if (!seen.emplace(key, 1).second) { is_dup[i] = 1; dup = true; }
                }
                if (dup) {
                    mdk_bam_batch *clean = new (std::nothrow) mdk_bam_batch();
                    if (!clean) { delete batch; set_error("bam_fetch: out of host memory"); return MDK_ERR_NOMEM; }
                    for (size_t i = 0; i < n; ++i) {
                        if (is_dup[i]) continue;
                        clean->pos.push_back(batch->pos[i]); clean->flag.push_back(batch->flag[i]);
                        clean->mapq.push_back(batch->mapq[i]); clean->l_seq.push_back(batch->l_seq[i]);
                        clean->cigar.insert(clean->cigar.end(), batch->cigar.begin() + batch->cigar_off[i], batch->cigar.begin() + batch->cigar_off[i + 1]);
                        clean->cigar_off.push_back((int64_t)clean->cigar.size());
                        clean->seq.insert(clean->seq.end(), batch->seq.begin() + batch->seq_off[i], batch->seq.begin() + batch->seq_off[i + 1]);
                        clean->seq_off.push_back((int64_t)clean->seq.size());
                        clean->qual.insert(clean->qual.end(), batch->qual.begin() + batch->qual_off[i], batch->qual.begin() + batch->qual_off[i + 1]);
                        clean->qual_off.push_back((int64_t)clean->qual.size());
                        clean->aux.insert(clean->aux.end(), batch->aux.begin() + batch->aux_off[i], batch->aux.begin() + batch->aux_off[i + 1]);
                        clean->aux_off.push_back((int64_t)clean->aux.size());
                        clean->names.insert(clean->names.end(), batch->names.begin() + batch->name_off[i], batch->names.begin() + batch->name_off[i + 1]);
                        clean->name_off.push_back((int64_t)clean->names.size());
                    }
                    delete batch;
                    batch = clean;
                }
            }
        } else {
            // no index: one bounded-memory pass over the file from the first record
            rc = stream_range(b, (int64_t)(b->first_voffset >> 16), b->file_size, (uint32_t)(b->first_voffset & 0xffff), tid,
                              start, end, exclude_flags, min_mapq, threads, batch, true);
        }
    }
    if (rc != MDK_OK) { delete batch; return rc; }
    *out = batch;
    return MDK_OK;
}

int64_t mdk_bam_batch_size(mdk_bam_batch *x) { return x ? (int64_t)x->pos.size() : 0; }

int mdk_bam_batch_arrays(mdk_bam_batch *x, const int32_t **pos, const uint16_t **flag, const uint8_t **mapq,
                         const int32_t **l_seq, const uint32_t **cigar, const int64_t **cigar_off, const uint8_t **seq,
                         const int64_t **seq_off, const uint8_t **aux, const int64_t **aux_off, const char **names,
                         const int64_t **name_off) {
    MDK_REQUIRE(x, MDK_ERR_ARG, "bam_batch_arrays: NULL batch");
    if (pos) *pos = x->pos.data();
    if (flag) *flag = x->flag.data();
    if (mapq) *mapq = x->mapq.data();
    if (l_seq) *l_seq = x->l_seq.data();
    if (cigar) *cigar = x->cigar.data();
    if (cigar_off) *cigar_off = x->cigar_off.data();
    if (seq) *seq = x->seq.data();
    if (seq_off) *seq_off = x->seq_off.data();
    if (aux) *aux = x->aux.data();
    if (aux_off) *aux_off = x->aux_off.data();
    if (names) *names = x->names.data();
    if (name_off) *name_off = x->name_off.data();
    return MDK_OK;
}

int mdk_bam_batch_qual(mdk_bam_batch *x, const uint8_t **qual, const int64_t **qual_off) {
    MDK_REQUIRE(x, MDK_ERR_ARG, "bam_batch_qual: NULL batch");
    if (qual) *qual = x->qual.data();
    if (qual_off) *qual_off = x->qual_off.data();
    return MDK_OK;
}

int mdk_bam_batch_free(mdk_bam_batch *x) {
    delete x;
    return MDK_OK;
}

}  // extern "C"
