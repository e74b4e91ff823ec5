#include "output.hpp"

#include <atomic>
#include <charconv>
#include <cmath>
#include <limits>
#include <thread>

namespace rgh {

namespace {

// "<label> v1 v2 ... " -> the label and the values (NA / nan / inf -> NaN, like convertDouble's missing code)
void parse_pred_row(const std::string& line, std::string& label, std::vector<double>& vals, size_t expect) {
  vals.clear();
  vals.reserve(expect);
  const char* p = line.c_str();
  const char* const end = p + line.size();
  auto skip = [&] { while (p < end && (*p == ' ' || *p == '\t')) ++p; };
  skip();
  const char* t0 = p;
  while (p < end && *p != ' ' && *p != '\t') ++p;
  label.assign(t0, (size_t)(p - t0));
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (;;) {
    skip();
    if (p >= end) break;
    if (p[0] == 'N' && p + 1 < end && p[1] == 'A' && (p + 2 == end || p[2] == ' ' || p[2] == '\t')) {
      vals.push_back(nan);
      p += 2;
      continue;
    }
    char* q = nullptr;
    const double v = std::strtod(p, &q);
    if (q == p) {
      const char* e = p;
      while (e < end && *e != ' ' && *e != '\t') ++e;
      throw Fail("could not convert value to double: '" + std::string(p, (size_t)(e - p)) + "'");
    }
    vals.push_back(std::isfinite(v) ? v : nan);
    p = q;
  }
}

}  // namespace

Loco read_loco(const std::string& path, bool prs) {
  LineReader fh(path);
  Loco l;
  l.prs = prs;
  l.path_ = path;
  l.eager_ = fh.is_gz();
  l.offs_.assign(23, -1);
  l.rows_.assign(23, {});
  std::string line, label;
  fh.getline(line);
  l.ids = split_ws(line);
  if (l.ids.empty() || l.ids[0] != "FID_IID")
    throw Fail("header of blup file must start with FID_IID" + (l.ids.empty() ? std::string(".") : " (=" + l.ids[0] + ")"));
  l.ids.erase(l.ids.begin());
  bool first = true;
  std::vector<double> vals;
  for (;;) {
    const int64_t off = fh.tell();
    if (!fh.getline(line)) break;
    size_t k = 0;
    while (k < line.size() && (line[k] == ' ' || line[k] == '\t')) ++k;
    if (k == line.size()) continue;
    size_t e = k;
    while (e < line.size() && line[e] != ' ' && line[e] != '\t') ++e;
    label.assign(line, k, e - k);
    if (prs) {                                               // src/Pheno.cpp:1297-1298
      if (!first) break;
      if (label != "0") throw Fail("second line must start with 0 (=" + label + ").");
    }
    const int c = prs ? 1 : chr_str_to_int(label);
    if (c < 1) throw Fail("blup file has an invalid chromosome row: " + label);
    if (first || l.eager_) {
      parse_pred_row(line, label, vals, l.ids.size());
      if (vals.size() != l.ids.size()) throw Fail("blup file has different number of entries compared to the header");
      if (first) l.first = vals;
      if (l.eager_) l.rows_[c - 1] = vals;
    }
    l.offs_[c - 1] = off;
    first = false;
  }
  return l;
}

bool Loco::has_row(int chrom) const {
  const int r = prs ? 0 : chrom - 1;
  return !empty() && r >= 0 && r < 23 && offs_[r] >= 0;
}

const std::vector<double>& Loco::row(int chrom) {
  const int r = prs ? 0 : chrom - 1;
  if (!has_row(chrom)) throw Fail("blup file " + path_ + " has no row for chromosome " + std::to_string(chrom));
  if (eager_ || cached_ == r) return rows_[r];
  if (cached_ >= 0) std::vector<double>().swap(rows_[cached_]);
  LineReader fh(path_);
  fh.seek(offs_[r]);
  std::string line, label;
  if (!fh.getline(line)) throw Fail("cannot read from file : " + path_);
  parse_pred_row(line, label, rows_[r], ids.size());
  if (rows_[r].size() != ids.size()) throw Fail("blup file has different number of entries compared to the header");
  cached_ = r;
  return rows_[r];
}

std::map<std::string, std::string> read_pred_list(const std::string& path) {
  std::map<std::string, std::string> blup_files;
  LineReader fh(path);
  std::string line;
  while (fh.getline(line)) {
    auto t = split_ws(line);
    if (t.empty()) continue;
    if (t.size() != 2) throw Fail("step 1 list file is not in the right format : " + path);
    if (blup_files.count(t[0])) throw Fail("phenotype '" + t[0] + "' appears more than once in step 1 list file.");
    blup_files[t[0]] = t[1];
  }
  return blup_files;
}

void write_pred_file(TextWriter& out, const std::vector<std::string>& keys, const std::vector<uint32_t>& order,
                     const uint8_t* mask, const std::vector<int>& row_labels, const std::vector<const double*>& values) {
  std::string buf;                                           // `ostream << double` == printf("%g") (6 significant digits)
  buf.reserve(order.size() * 12 + 64);
  buf += "FID_IID ";
  for (uint32_t i : order) { buf += keys[i]; buf += ' '; }
  buf += '\n';
  out << buf;
  // rows and sample ranges are independent: format (row, chunk of samples) items on the host threads (23 x N numbers; at
  // N = 500k this is the longest host-side step of Step 1 when done serially), write them in order.  std::to_chars with
  // chars_format::general and precision 6 is specified to give printf("%g")'s characters, i.e. what `ostream << double`
  // of the reference prints (src/Data.cpp:1836-1870), at less than half the cost of snprintf.
  const size_t R = row_labels.size(), n = order.size();
  const size_t chunk = 32768, nchunk = std::max<size_t>(1, (n + chunk - 1) / chunk), items = R * nchunk;
  std::vector<std::string> bufs(items);
  auto format_item = [&](size_t it) {
    const size_t r = it / nchunk, c = it % nchunk, lo = c * chunk, hi = std::min(n, lo + chunk);
    std::string& b = bufs[it];
    b.reserve((hi - lo) * 10 + 16);
    if (c == 0) {
      b += std::to_string(row_labels[r]);
      b += ' ';
    }
    const double* v = values[r];
    char num[48];
    for (size_t k = lo; k < hi; ++k) {
      const uint32_t i = order[k];
      if (mask[i]) {
        char* e = std::to_chars(num, num + 40, v[i], std::chars_format::general, 6).ptr;
        *e++ = ' ';
        b.append(num, (size_t)(e - num));
      } else b += "NA ";
    }
    if (c + 1 == nchunk) b += '\n';
  };
  const size_t work_items = R * n;
  size_t T = work_items > (1u << 18) ? std::min<size_t>({items, 32, std::max(1u, std::thread::hardware_concurrency())}) : 1;
  std::atomic<size_t> next{0};
  auto worker = [&] {
    for (;;) {
      const size_t it = next.fetch_add(1);
      if (it >= items) return;
      format_item(it);
    }
  };
  std::vector<std::thread> pool;
  for (size_t t = 1; t < T; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  for (size_t it = 0; it < items; ++it) out << bufs[it];
}

double get_logp(double t) {   // chi2_1 sf = erfc(sqrt(T/2))
  if (t < 0 && std::fabs(t) < 1e-6) return 0.0;
  if (t < 0) return -1.0;
  const double pv = std::erfc(std::sqrt(t / 2.0));
  double lp;
  if (pv == 0) lp = std::log10(2.0) - 0.5 * std::log10(2 * M_PI * t) - 0.5 * t * M_LOG10E;
  else lp = std::log10(pv);
  return -lp;
}

std::string sumstats_header(bool with_info, bool af_cc) {
  return std::string("CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ ") + (af_cc ? "A1FREQ_CASES A1FREQ_CONTROLS " : "") +
         (with_info ? "INFO " : "") + "N " + (af_cc ? "N_CASES N_CONTROLS " : "") + "TEST BETA SE CHISQ LOG10P EXTRA\n";
}

// `ostream << double` at the default precision prints like printf("%g"); snprintf into the caller's buffer is several
// times cheaper than a stringstream per row, which matters at 10^7 variants x P traits
void append_sumstats_row(std::string& out, const std::string& head, double af, bool with_info, double info, int n,
                         const char* test, double beta, double se, double chisq, double logp, bool test_pass,
                         const AfCc* cc) {
  char num[96];
  out += head;
  if (af >= 0) out.append(num, (size_t)snprintf(num, sizeof(num), "%g ", af));
  else out += "NA ";
  if (cc) {
    if (af >= 0) out.append(num, (size_t)snprintf(num, sizeof(num), "%g %g ", cc->af_case, cc->af_control));
    else out += "NA NA ";
  }
  if (with_info) {
    if (info >= 0) out.append(num, (size_t)snprintf(num, sizeof(num), "%g ", info));
    else out += "NA ";
  }
  out.append(num, (size_t)snprintf(num, sizeof(num), "%d ", n));
  if (cc) out.append(num, (size_t)snprintf(num, sizeof(num), "%d %d ", cc->ns_case, cc->ns_control));
  out += test;
  out += ' ';
  if (se >= 0 && !std::isnan(se)) out.append(num, (size_t)snprintf(num, sizeof(num), "%g %g", beta, se));
  else out += "NA NA";
  if (chisq >= 0 && test_pass && !std::isnan(logp)) out.append(num, (size_t)snprintf(num, sizeof(num), " %g %g", chisq, logp));
  else out += " NA NA";
  out += test_pass ? " NA\n" : " TEST_FAIL\n";
}

std::string sumstats_row(const std::string& head, double af, bool with_info, double info, int n, const char* test,
                         double beta, double se, double chisq, double logp, bool test_pass) {
  std::string out;
  append_sumstats_row(out, head, af, with_info, info, n, test, beta, se, chisq, logp, test_pass);
  return out;
}

std::string sumstats_header_all(int n_pheno, bool with_info) {
  std::string h = std::string("CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ ") + (with_info ? "INFO " : "") + "N N_RR N_RA N_AA TEST";
  for (int i = 1; i <= n_pheno; ++i) {
    const std::string k = std::to_string(i);
    h += " BETA.Y" + k + " SE.Y" + k + " CHISQ.Y" + k + " LOG10P.Y" + k;
  }
  return h + " EXTRA\n";
}

void append_sumstats_all_start(std::string& out, const std::string& head, double af, int n, long n_rr, long n_ra, long n_aa,
                               const char* test, bool with_info, double info) {
  char num[128];
  out += head;
  if (af >= 0) out.append(num, (size_t)snprintf(num, sizeof(num), "%g", af));
  else out += "NA";
  if (with_info) {
    if (info >= 0) out.append(num, (size_t)snprintf(num, sizeof(num), " %g", info));
    else out += " NA";
  }
  out.append(num, (size_t)snprintf(num, sizeof(num), " %d", n));
  if (n_rr >= 0) out.append(num, (size_t)snprintf(num, sizeof(num), " %ld %ld %ld", n_rr, n_ra, n_aa));
  else out += " NA NA NA";
  out += ' ';
  out += test;
}

void append_sumstats_all_trait(std::string& out, bool have, double beta, double se, double chisq, double logp, bool test_pass) {
  char num[96];
  if (have && se >= 0 && !std::isnan(se)) out.append(num, (size_t)snprintf(num, sizeof(num), " %g %g", beta, se));
  else out += " NA NA";
  if (have && chisq >= 0 && test_pass && !std::isnan(logp)) out.append(num, (size_t)snprintf(num, sizeof(num), " %g %g", chisq, logp));
  else out += " NA NA";
}

std::string htp_header() {
  return "Name\tChr\tPos\tRef\tAlt\tTrait\tCohort\tModel\tEffect\tLCI_Effect\tUCI_Effect\tPval\tAAF\tNum_Cases\tCases_Ref\tCases_Het\t"
         "Cases_Alt\tNum_Controls\tControls_Ref\tControls_Het\tControls_Alt\tInfo\n";
}

static std::string dbl_to_str(double v) {                      // convert_double_to_str, src/Regenie.cpp:1691-1698
  char b[64];
  if (v < 5000 && v > 1e-5) snprintf(b, sizeof(b), "%.6f", v);
  else snprintf(b, sizeof(b), "%g", v);
  return b;
}

static std::string logp_raw(double logp) {                     // convert_logp_raw, src/Regenie.cpp:1700-1717
  char b[64];
  const double log_dbl_min = -std::log10(std::numeric_limits<double>::min()) - 1;
  if (logp <= 3) snprintf(b, sizeof(b), "%f", std::pow(10, -logp));
  else if (logp <= log_dbl_min) snprintf(b, sizeof(b), "%g", std::pow(10, -logp));
  else {
    const double thr = std::log(9.95) / std::log(10);
    int base = (int)std::ceil(logp);
    double res = base - logp;
    if (res >= thr) { res = 0; base++; }
    snprintf(b, sizeof(b), "%.1fe-%d", std::pow(10, res), base);
  }
  return b;
}

void append_htp_row(std::string& out, const std::string& head, const std::string& trait, const std::string& cohort, const HtpRow& r) {
  static const double zcrit = 1.959963984540054;              // quantile(complement(normal, .025)), src/Data.cpp:2118
  static const double log10_nl_dbl_dmin = -std::log10(10.0 * std::numeric_limits<double>::min());   // src/Regenie.hpp:229-230
  char num[160];
  auto g = [&](double v) { out.append(num, (size_t)snprintf(num, sizeof(num), "%g", v)); };
  out += head; out += trait; out += '\t'; out += cohort; out += '\t'; out += r.model; out += '\t';
  const bool print_beta = r.test_pass && r.se >= 0 && !std::isnan(r.se);
  const bool print_pv = r.test_pass && r.chisq >= 0 && !std::isnan(r.logp);
  std::string outp = "-1";
  if (print_pv) {
    if (r.logp > log10_nl_dbl_dmin) outp = logp_raw(log10_nl_dbl_dmin);
    else if (r.logp > 0) outp = logp_raw(r.logp);
    else outp = "0.9999999";
  }
  double outse = 0;
  if (print_pv && !print_beta) { out += "NA\tNA\tNA\t"; out += outp; out += '\t'; }
  else if (!print_pv && !print_beta) out += "NA\tNA\tNA\tNA\t";
  else if (!r.bt || (r.bt && r.firth && r.test_pass)) {
    if (!r.bt) { g(r.beta); out += '\t'; g(r.beta - zcrit * r.se); out += '\t'; g(r.beta + zcrit * r.se); out += '\t'; }
    else { g(std::exp(r.beta)); out += '\t'; g(std::exp(r.beta - zcrit * r.se)); out += '\t'; g(std::exp(r.beta + zcrit * r.se)); out += '\t'; }
    out += print_pv ? outp : "NA";
    out += '\t';
  } else if (print_pv) {                                       // SPA / uncorrected logistic score test: allelic odds ratio
    const double eff = (2 * r.gc[3] + r.gc[4] + .5) * (2 * r.gc[2] + r.gc[1] + .5) / (2 * r.gc[5] + r.gc[4] + .5) / (2 * r.gc[0] + r.gc[1] + .5);
    outse = std::fabs(std::log(eff)) / std::sqrt(r.chisq);
    g(eff); out += '\t'; g(eff * std::exp(-zcrit * outse)); out += '\t'; g(eff * std::exp(zcrit * outse)); out += '\t';
    out += outp; out += '\t';
  } else {
    g(std::exp(r.beta)); out += '\t'; g(std::exp(r.beta - zcrit * r.se)); out += '\t'; g(std::exp(r.beta + zcrit * r.se)); out += "\tNA\t";
  }
  if (r.af >= 0) { g(r.af); out += '\t'; } else out += "NA\t";
  out.append(num, (size_t)snprintf(num, sizeof(num), "%ld\t%ld\t%ld\t%ld\t", r.gc[0] + r.gc[1] + r.gc[2], r.gc[0], r.gc[1], r.gc[2]));
  if (r.bt) out.append(num, (size_t)snprintf(num, sizeof(num), "%ld\t%ld\t%ld\t%ld", r.gc[3] + r.gc[4] + r.gc[5], r.gc[3], r.gc[4], r.gc[5]));
  else out += "NA\tNA\tNA\tNA";
  std::string col;
  auto add = [&](const std::string& t) { if (!col.empty()) col += ';'; col += t; };
  if (print_beta) {
    if (r.bt && r.test_pass) {
      add("REGENIE_BETA=" + dbl_to_str(r.beta));
      add("REGENIE_SE=" + dbl_to_str(r.se));
      if (print_pv && !r.firth) add("SE=" + dbl_to_str(outse));
    } else if (r.bt) {
      add("REGENIE_BETA=NA");
      add("REGENIE_SE=NA");
    } else add("REGENIE_SE=" + std::to_string(r.se));
  }
  if (r.info >= 0) add("INFO=" + dbl_to_str(r.info));
  if (r.mac >= 0) add("MAC=" + std::to_string(r.mac));
  if (r.has_score) {
    add("SCORE=" + dbl_to_str(r.score));
    add("SKATV=" + dbl_to_str(r.skat_var * std::fabs(r.cal_factor)));
  }
  add("LOG10P=" + (print_pv ? dbl_to_str(r.logp) : std::string("NA")));
  if (r.se < 0) add("NO_BETA");
  out += '\t';
  out += col;
  out += '\n';
}

void write_ids_file(const std::string& path, const std::string& pheno_name, bool print_pheno_name,
                    const std::vector<std::pair<std::string, std::string>>& fid_iid, const uint8_t* mask) {
  TextWriter out;
  out.open(path);
  std::string buf;
  if (print_pheno_name) buf += pheno_name + "\tNA\n";
  bool first = true;
  for (size_t i = 0; i < fid_iid.size(); ++i) {
    if (!mask[i]) continue;
    if (!first) buf += '\n';
    first = false;
    buf += fid_iid[i].first;
    buf += '\t';
    buf += fid_iid[i].second;
  }
  out << buf;
  out.close();
}

}  // namespace rgh
