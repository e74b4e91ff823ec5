#include "pgen.hpp"
#include "textio.hpp"

#include <cstring>

namespace rgh {

namespace {
constexpr uint32_t kVblock = 65536, kGroup = 64;

inline uint32_t vint(const std::vector<uint8_t>& d, size_t& p) {
  uint32_t v = 0, shift = 0;
  for (;;) {
    if (p >= d.size()) throw Fail("malformed .pgen file (variable-length integer runs past the end).");
    const uint8_t b = d[p++];
    if (shift > 28) throw Fail("malformed .pgen file (variable-length integer is too long).");
    v |= (uint32_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) return v;
    shift += 7;
  }
}
}  // namespace

void PgenFile::open(const std::string& pfx, const std::set<std::string>& exclude, const std::set<std::string>& extract,
                    const std::set<std::string>& remove, const std::set<std::string>& keep, const std::set<int>& chrs) {
  prefix = pfx;
  // ---- .psam (src/Geno.cpp:941-1011)
  {
    LineReader fh(prefix + ".psam");
    std::string line;
    int sex_col = -1;
    bool header = true;
    std::set<std::string> seen;
    while (fh.getline(line)) {
      auto t = split_ws(line);
      if (t.empty()) continue;
      if (header) {
        if (t.size() < 2 || t[0] != "#FID" || t[1] != "IID") throw Fail("header does not have the correct format (must start with #FID IID).");
        for (size_t i = 2; i < t.size(); ++i) if (t[i] == "SEX") sex_col = (int)i;
        header = false;
        continue;
      }
      if (t.size() < 2) throw Fail("incorrectly formatted psam file.");
      const std::string k = t[0] + "_" + t[1];
      if (!seen.insert(k).second) throw Fail("duplicate individual in psam file : FID_IID=" + k);
      keys_file.push_back(k);
      ids_file.emplace_back(t[0], t[1]);
      int sx = 0;
      if (sex_col >= 0 && (size_t)sex_col < t.size()) sx = t[sex_col] == "1" ? 1 : (t[sex_col] == "2" ? 2 : 0);
      sex_file.push_back(sx);
    }
  }
  for (size_t i = 0; i < keys_file.size(); ++i) {
    const std::string& k = keys_file[i];
    if (remove.count(k)) continue;
    if (!keep.empty() && !keep.count(k)) continue;
    if (sex_specific && sex_file[i] != sex_specific) continue;   // --sex-specific (src/Geno.cpp:1287-1293)
    key_to_ind[k] = (uint32_t)keys.size();
    keys.push_back(k);
    sample_idx.push_back((int32_t)i);
  }
  if (keys.empty()) throw Fail("no samples left after --keep/--remove/--sex-specific.");
  // ---- .pvar (src/Geno.cpp:771-870): ALLELE0 = REF, ALLELE1 = ALT
  {
    LineReader fh(prefix + ".pvar");
    std::string line;
    int ipos = -1, iid = -1, iref = -1, ialt = -1;
    uint64_t idx = 0;
    int last_chr = 0;
    std::vector<int> chr_seen;
    while (fh.getline(line)) {
      if (line.rfind("##", 0) == 0) continue;
      auto t = split_ws(line);
      if (t.empty()) continue;
      if (ipos < 0) {
        if (t[0] != "#CHROM") throw Fail("header of pvar file does not have correct format.");
        for (size_t i = 0; i < t.size(); ++i) {
          if (t[i] == "POS") ipos = (int)i; else if (t[i] == "ID") iid = (int)i;
          else if (t[i] == "REF") iref = (int)i; else if (t[i] == "ALT") ialt = (int)i;
        }
        if (ipos < 0 || iid < 0 || iref < 0 || ialt < 0) throw Fail("header of pvar file does not have correct format.");
        continue;
      }
      if (t.size() < 5) throw Fail("incorrectly formatted pvar file at line " + std::to_string(idx + 1));
      Snp s;
      s.chrom = chr_str_to_int(t[0]);
      if (s.chrom == -1) throw Fail("unknown chromosome code in pvar file at line " + std::to_string(idx + 1));
      if (chr_seen.empty() || chr_seen.back() != s.chrom) {
        if (s.chrom <= last_chr) throw Fail("chromosomes in pvar file are not in ascending order.");
        chr_seen.push_back(s.chrom);
        last_chr = s.chrom;
      }
      s.pos = std::stoull(t[ipos], nullptr, 0);
      s.id = t[iid];
      s.allele0 = t[iref];
      s.allele1 = t[ialt];
      if (s.allele1.find(',') != std::string::npos) throw Fail("only bi-allelic variants are supported in pgen files (variant " + s.id + ").");
      s.offset = idx++;
      if (!chrs.empty() && !chrs.count(s.chrom)) continue;
      if (exclude.count(s.id)) continue;
      if (!extract.empty() && !extract.count(s.id)) continue;
      snps.push_back(s);
    }
    m_file = (uint32_t)idx;
  }
  // ---- .pgen header (PgfiInitPhase1/2)
  {
    std::ifstream fh(prefix + ".pgen", std::ios::binary | std::ios::ate);
    if (!fh) throw Fail("cannot open file : " + prefix + ".pgen");
    const std::streamsize sz = fh.tellg();
    fh.seekg(0);
    data.resize((size_t)sz);
    fh.read(reinterpret_cast<char*>(data.data()), sz);
    if (!fh) throw Fail("cannot read from pgen file.");
  }
  const std::vector<uint8_t>& d = data;
  if (d.size() < 12 || d[0] != 0x6c || d[1] != 0x1b) throw Fail("invalid pgen file (magic number).");
  const uint8_t mode = d[2];
  uint32_t m = 0, n = 0;
  memcpy(&m, &d[3], 4); memcpy(&n, &d[7], 4);
  const uint8_t ctrl = d[11];
  if (m != m_file) throw Fail("number of variants in the .pgen file does not match the .pvar file.");
  if (n != keys_file.size()) throw Fail("number of samples in the .pgen file does not match the .psam file.");
  n_file = n;
  row_stride = ((uint64_t)n + 3) / 4;
  const uint64_t n4 = row_stride;
  vrtype.assign(m, 0);
  fpos.assign((size_t)m + 1, 0);
  if (mode == 0x02) {
    if (ctrl & 63) throw Fail("invalid pgen file (fixed-width mode with a variable-width header byte).");
    const uint64_t off = 12 + ((ctrl >> 6) == 3 ? ((uint64_t)m + 7) / 8 : 0);
    for (uint32_t v = 0; v <= m; ++v) fpos[v] = off + n4 * v;
    if (fpos[m] != d.size()) throw Fail("unexpected .pgen file size.");
  } else if (mode == 0x10) {
    const uint32_t storage = ctrl & 15;
    if (storage >= 8) throw Fail("this .pgen header layout (single-sample fused record types) is not supported by rgb200.");
    if ((ctrl >> 4) & 3) throw Fail("multiallelic .pgen files are not supported (split them with plink2 first).");
    const bool nonref_stored = (ctrl >> 6) == 3;
    const uint32_t nblk = (m - 1) / kVblock + 1;
    size_t p = 12;
    auto need = [&](size_t k) { if (p + k > d.size()) throw Fail("malformed .pgen header."); };
    need((size_t)8 * nblk);
    std::vector<uint64_t> blk(nblk);
    for (uint32_t b = 0; b < nblk; ++b) memcpy(&blk[b], &d[p + 8 * b], 8);
    p += (size_t)8 * nblk;
    const uint32_t lb = 1 + (storage & 3);
    uint64_t cur = 0;
    for (uint32_t b = 0; b < nblk; ++b) {
      const uint32_t cnt = std::min<uint32_t>(kVblock, m - b * kVblock), v0 = b * kVblock;
      if (storage < 4) {
        need((cnt + 1) / 2);
        for (uint32_t i = 0; i < cnt; ++i) vrtype[v0 + i] = (d[p + i / 2] >> (4 * (i & 1))) & 15;
        p += (cnt + 1) / 2;
      } else {
        need(cnt);
        memcpy(&vrtype[v0], &d[p], cnt);
        p += cnt;
      }
      need((size_t)cnt * lb);
      cur = blk[b];
      for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t len = 0;
        memcpy(&len, &d[p + (size_t)i * lb], lb);
        fpos[v0 + i] = cur;
        cur += len;
      }
      p += (size_t)cnt * lb;
      if (nonref_stored) { need((cnt + 7) / 8); p += (cnt + 7) / 8; }
    }
    fpos[m] = cur;
    if (cur > d.size()) throw Fail("malformed .pgen file (records run past the end).");
  } else {
    throw Fail("this .pgen storage mode (dosages / PLINK 1 bed / extensions) is not supported by rgb200; hard-call .pgen only.");
  }
  for (uint32_t v = 0; v < m; ++v)
    if (vrtype[v] & 0xE8) throw Fail("the .pgen file holds multiallelic or dosage tracks: not supported by rgb200 (hard calls only).");
  base_.assign(n, 0);
  cur_.assign(n, 0);
}

size_t PgenFile::difflist(size_t p, std::vector<uint32_t>& ids, std::vector<uint8_t>& vals) const {
  const std::vector<uint8_t>& d = data;
  const uint32_t len = vint(d, p);
  ids.resize(len); vals.resize(len);
  if (!len) return p;
  if (len > n_file) throw Fail("malformed .pgen record (difflist longer than the sample count).");
  const uint32_t ng = (len + kGroup - 1) / kGroup;
  const uint32_t sb = n_file <= 0xFF ? 1 : n_file <= 0xFFFF ? 2 : n_file <= 0xFFFFFF ? 3 : 4;
  const size_t first = p;
  p += (size_t)ng * (sb + 1) - 1;
  const size_t nv = (len + 3) / 4;
  if (p + nv > d.size()) throw Fail("malformed .pgen record.");
  for (uint32_t i = 0; i < len; ++i) vals[i] = (d[p + i / 4] >> (2 * (i & 3))) & 3;
  p += nv;
  uint32_t k = 0;
  for (uint32_t g = 0; g < ng; ++g) {
    uint32_t cur = 0;
    memcpy(&cur, &d[first + (size_t)g * sb], sb);
    ids[k++] = cur;
    const uint32_t cnt = std::min<uint32_t>(kGroup, len - g * kGroup);
    for (uint32_t j = 1; j < cnt; ++j) { cur += vint(d, p); ids[k++] = cur; }
    if (cur >= n_file) throw Fail("malformed .pgen record (sample index out of range).");
  }
  return p;
}

void PgenFile::decode_nonld(uint32_t v, std::vector<uint8_t>& g) const {
  const std::vector<uint8_t>& d = data;
  const uint32_t t = vrtype[v] & 7, n = n_file;
  size_t p = fpos[v];
  std::vector<uint32_t> ids;
  std::vector<uint8_t> vals;
  if (t == 0) {                                   // plain 2-bit
    if (p + row_stride > d.size()) throw Fail("malformed .pgen record.");
    for (uint32_t i = 0; i < n; ++i) g[i] = (d[p + i / 4] >> (2 * (i & 3))) & 3;
    return;
  }
  if (t == 1) {                                   // 1 bit per sample between two values, exceptions in a difflist
    const size_t nb = ((size_t)n + 7) / 8;
    if (p + 1 + nb > d.size()) throw Fail("malformed .pgen record.");
    const uint8_t code = d[p], lo = code >> 2, delta = code & 3;
    for (uint32_t i = 0; i < n; ++i) g[i] = lo + ((d[p + 1 + i / 8] >> (i & 7)) & 1) * delta;
    difflist(p + 1 + nb, ids, vals);
  } else if (t == 5) {                            // all homozygous reference
    std::fill(g.begin(), g.end(), 0);
    return;
  } else {                                        // 4 / 6 / 7: constant 0 / 2 / missing with a difflist
    std::fill(g.begin(), g.end(), (uint8_t)(t & 3));
    difflist(p, ids, vals);
  }
  for (size_t k = 0; k < ids.size(); ++k) g[ids[k]] = vals[k];
}

void PgenFile::decode(uint32_t v) {
  const uint32_t t = vrtype[v] & 7;
  if ((t & 6) != 2) {
    decode_nonld(v, cur_);
    base_ = cur_;
    base_idx_ = v;
    return;
  }
  long b = (long)v - 1;                           // LD-compressed: difflist against the latest non-LD record
  while (b >= 0 && (vrtype[b] & 6) == 2) --b;
  if (b < 0) throw Fail("malformed .pgen file (LD-compressed record without a base).");
  if (base_idx_ != b) { decode_nonld((uint32_t)b, base_); base_idx_ = b; }
  cur_ = base_;
  std::vector<uint32_t> ids;
  std::vector<uint8_t> vals;
  difflist(fpos[v], ids, vals);
  for (size_t k = 0; k < ids.size(); ++k) cur_[ids[k]] = vals[k];
  if (t == 3)
    for (auto& x : cur_) x = (x == 0) ? 2 : (x == 2 ? 0 : x);
}

void PgenFile::read_rows(size_t first, size_t n, uint8_t* out) {
  static const uint8_t kBed[4] = {3, 2, 0, 1};    // ALT count 0 / 1 / 2 / missing -> PLINK 1 code (ref-last)
  for (size_t j = 0; j < n; ++j) {
    decode((uint32_t)snps[first + j].offset);
    uint8_t* row = out + j * row_stride;
    memset(row, 0, row_stride);
    for (uint32_t i = 0; i < n_file; ++i) {
      if (cur_[i] > 3) throw Fail("malformed .pgen record (genotype value out of range).");
      row[i >> 2] |= (uint8_t)(kBed[cur_[i]] << (2 * (i & 3)));
    }
  }
}

void PgenFile::gather(size_t first, size_t n, PgenBatch& out) const {
  out.bytes.clear(); out.rec_off.clear(); out.rec_len.clear(); out.rec_type.clear();
  out.own.assign(n, -1); out.base.assign(n, -1);
  long last_base_v = -1;
  int32_t last_base_rec = -1;
  auto add = [&](uint32_t v) -> int32_t {
    const uint64_t len = fpos[v + 1] - fpos[v];
    if (fpos[v + 1] < fpos[v] || fpos[v + 1] > data.size() || len >= (1ull << 32)) throw Fail("malformed .pgen file (record offsets).");
    const size_t off = (out.bytes.size() + 15) / 16 * 16;
    out.bytes.resize(off + len, 0);
    if (len) memcpy(&out.bytes[off], &data[fpos[v]], len);
    out.rec_off.push_back(off);
    out.rec_len.push_back((uint32_t)len);
    out.rec_type.push_back(vrtype[v] & 7);
    return (int32_t)out.rec_off.size() - 1;
  };
  for (size_t j = 0; j < n; ++j) {
    const uint32_t v = (uint32_t)snps[first + j].offset;
    const uint32_t t = vrtype[v] & 7;
    out.own[j] = add(v);
    if ((t & 6) != 2) {                          // a later LD record of the block may refer to this one
      last_base_v = v;
      last_base_rec = out.own[j];
      continue;
    }
    long b = (long)v - 1;
    while (b >= 0 && (vrtype[b] & 6) == 2) --b;
    if (b < 0) throw Fail("malformed .pgen file (LD-compressed record without a base).");
    if (b != last_base_v) { last_base_rec = add((uint32_t)b); last_base_v = b; }
    out.base[j] = last_base_rec;
  }
  if (out.bytes.empty()) out.bytes.resize(16, 0);           // a block of all-reference records has no bytes at all
}

void pgen_read_rows(PgenFile& pg, size_t first, size_t n, uint8_t* out) { pg.read_rows(first, n, out); }

void BedFile::open_pgen(const std::string& pfx, const std::set<std::string>& exclude, const std::set<std::string>& extract,
                        const std::set<std::string>& remove, const std::set<std::string>& keep, const std::set<int>& chrs) {
  pg = std::make_shared<PgenFile>();
  pg->sex_specific = sex_specific;
  pg->open(pfx, exclude, extract, remove, keep, chrs);
  prefix = pfx;
  snps = pg->snps; keys_file = pg->keys_file; ids_file = pg->ids_file; sex_file = pg->sex_file; keys = pg->keys;
  sample_idx = pg->sample_idx; key_to_ind = pg->key_to_ind; row_stride = pg->row_stride;
}

}  // namespace rgh
