// Small host utilities for the rgb200 driver: logging tee, string splitting, number formatting.
// (reference counterparts: mstream src/Regenie.hpp:120-143, string_split src/Files.cpp)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace rgh {

struct Fail : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// tee to stdout and <out>.log like the reference's `sout`
class Log {
 public:
  void open(const std::string& path) { file_.open(path); }
  void close() {
    std::cout.flush();
    if (file_.is_open()) file_.close();
  }
  template <typename T>
  Log& operator<<(const T& v) {
    std::cout << v;
    if (file_.is_open()) file_ << v;
    return *this;
  }
  Log& operator<<(std::ostream& (*m)(std::ostream&)) {
    std::cout << m;
    if (file_.is_open()) file_ << m;
    return *this;
  }

 private:
  std::ofstream file_;
};

inline std::vector<std::string> split_ws(const std::string& line) {
  std::vector<std::string> out;
  size_t i = 0, n = line.size();
  while (i < n) {
    while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r' || line[i] == '\n')) ++i;
    size_t j = i;
    while (j < n && !(line[j] == ' ' || line[j] == '\t' || line[j] == '\r' || line[j] == '\n')) ++j;
    if (j > i) out.emplace_back(line.substr(i, j - i));
    i = j;
  }
  return out;
}

// `ostream << double` at the default precision (6 significant digits), as the reference prints
inline std::string fmt_g(double v) {
  std::ostringstream b;
  b << v;
  return b.str();
}

constexpr double kMissing = -999.0;   // params.missing_value_double, src/Regenie.hpp:215

// convertDouble, src/Regenie.cpp:1663-1675
inline double convert_double(const std::string& s) {
  if (s == "NA" || s == "nan" || s == "inf") return kMissing;
  char* end = nullptr;
  const double v = std::strtod(s.c_str(), &end);
  if (end == s.c_str()) throw Fail("could not convert value to double: '" + s + "'");
  return v;
}

// chrStrToInt, src/Regenie.cpp:1583-1594
inline int chr_str_to_int(std::string s, int nchrom = 23) {
  if (s.rfind("chr", 0) == 0) s = s.substr(3);
  if (!s.empty() && isdigit((unsigned char)s[0])) {
    const int c = atoi(s.c_str());
    if (c >= 1 && c <= nchrom) return c;
  } else if (s == "X" || s == "XY" || s == "Y" || s == "PAR1" || s == "PAR2") {
    return nchrom;
  }
  return -1;
}

}  // namespace rgh
