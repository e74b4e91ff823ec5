#include "output.hpp"

#include <cmath>

namespace rgh {

Loco read_loco(const std::string& path, bool prs) {
  LineReader fh(path);
  Loco l;
  l.prs = prs;
  std::string line;
  fh.getline(line);
  l.ids = split_ws(line);
  if (l.ids.empty() || l.ids[0] != "FID_IID")
    throw Fail("header of blup file must start with FID_IID" + (l.ids.empty() ? std::string(".") : " (=" + l.ids[0] + ")"));
  l.ids.erase(l.ids.begin());
  l.rows.resize(23);
  bool first = true;
  while (fh.getline(line)) {
    auto t = split_ws(line);
    if (t.empty()) continue;
    if (prs) {                                               // src/Pheno.cpp:1297-1298
      if (!first) break;
      if (t[0] != "0") throw Fail("second line must start with 0 (=" + t[0] + ").");
    }
    first = false;
    const int c = prs ? 1 : chr_str_to_int(t[0]);
    if (c < 1) throw Fail("blup file has an invalid chromosome row: " + t[0]);
    if (t.size() != l.ids.size() + 1) throw Fail("blup file has different number of entries compared to the header");
    t.erase(t.begin());
    l.rows[c - 1] = std::move(t);
  }
  return l;
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
  char num[40];
  for (size_t r = 0; r < row_labels.size(); ++r) {
    buf.clear();
    buf += std::to_string(row_labels[r]);
    buf += ' ';
    const double* v = values[r];
    for (uint32_t i : order) {
      if (mask[i]) buf.append(num, (size_t)snprintf(num, sizeof(num), "%g ", v[i]));
      else buf += "NA ";
    }
    buf += '\n';
    out << buf;
  }
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

std::string sumstats_header(bool with_info) {
  return std::string("CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ ") + (with_info ? "INFO " : "") +
         "N TEST BETA SE CHISQ LOG10P EXTRA\n";
}

std::string sumstats_row(const std::string& head, double af, bool with_info, double info, int n, const char* test,
                         double beta, double se, double chisq, double logp, bool test_pass) {
  std::ostringstream buf;
  buf << head;
  if (af >= 0) buf << af << " ";
  else buf << "NA ";
  if (with_info) {
    if (info >= 0) buf << info << " ";
    else buf << "NA ";
  }
  buf << n << " " << test << " ";
  if (se >= 0 && !std::isnan(se)) buf << beta << ' ' << se;
  else buf << "NA NA";
  if (chisq >= 0 && test_pass && !std::isnan(logp)) buf << ' ' << chisq << ' ' << logp;
  else buf << " NA NA";
  buf << (test_pass ? " NA\n" : " TEST_FAIL\n");
  return buf.str();
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
