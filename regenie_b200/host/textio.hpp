// Text files of the rgb200 host driver, plain or gzip-compressed.  Like the reference's `Files` class
// (src/Files.cpp:39-150) the compression is decided by the file name alone: a name ending in ".gz" is read / written
// through zlib, anything else as plain text.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cstring>

#include "util.hpp"

namespace rgh {

inline bool has_gz_ext(const std::string& path) {
  return path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
}

// line-by-line reader (lines of any length: a .loco row holds one token per sample)
class LineReader {
 public:
  explicit LineReader(const std::string& path) : path_(path), buf_(1 << 16) {
    if (has_gz_ext(path)) {
      gz_ = gzopen(path.c_str(), "rb");
      if (gz_) gzbuffer(gz_, 1 << 18);
    } else {
      fp_ = fopen(path.c_str(), "rb");
    }
    if (!gz_ && !fp_) throw Fail("cannot open file : " + path);
  }
  LineReader(const LineReader&) = delete;
  LineReader& operator=(const LineReader&) = delete;
  ~LineReader() {
    if (gz_) gzclose(gz_);
    if (fp_) fclose(fp_);
  }
  bool is_gz() const { return gz_ != nullptr; }
  // byte offset (in the uncompressed text) of the next line getline() will return
  int64_t tell() const { return consumed_; }
  // plain files only: continue reading at byte offset `off`
  void seek(int64_t off) {
    if (!fp_ || fseeko(fp_, (off_t)off, SEEK_SET) != 0) throw Fail("cannot seek in file : " + path_);
    pos_ = len_ = 0;
    eof_ = false;
    consumed_ = off;
  }
  // false at end of file; the trailing "\n" / "\r\n" is removed
  bool getline(std::string& line) {
    line.clear();
    bool any = false;
    for (;;) {
      if (pos_ == len_) {
        if (eof_) break;
        fill();
        if (len_ == 0) break;
      }
      const char* b = buf_.data() + pos_;
      const char* nl = static_cast<const char*>(memchr(b, '\n', len_ - pos_));
      any = true;
      if (nl) {
        line.append(b, (size_t)(nl - b));
        consumed_ += (int64_t)(nl - b) + 1;
        pos_ += (size_t)(nl - b) + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return true;
      }
      line.append(b, len_ - pos_);
      consumed_ += (int64_t)(len_ - pos_);
      pos_ = len_;
    }
    if (any && !line.empty() && line.back() == '\r') line.pop_back();
    return any;
  }

 private:
  void fill() {
    pos_ = 0;
    if (gz_) {
      const int n = gzread(gz_, buf_.data(), (unsigned)buf_.size());
      if (n < 0) throw Fail("cannot read from file : " + path_);
      len_ = (size_t)n;
    } else {
      len_ = fread(buf_.data(), 1, buf_.size(), fp_);
      if (len_ == 0 && ferror(fp_)) throw Fail("cannot read from file : " + path_);
    }
    if (len_ == 0) eof_ = true;
  }
  std::string path_;
  std::vector<char> buf_;
  size_t pos_ = 0, len_ = 0;
  int64_t consumed_ = 0;
  bool eof_ = false;
  gzFile gz_ = nullptr;
  FILE* fp_ = nullptr;
};

// sequential writer; `path` ending in ".gz" selects gzip (callers append the suffix when --gz is set)
class TextWriter {
 public:
  TextWriter() = default;
  TextWriter(const TextWriter&) = delete;
  TextWriter& operator=(const TextWriter&) = delete;
  ~TextWriter() {                               // errors surface through an explicit close()
    if (gz_) gzclose(gz_);
    if (fp_) fclose(fp_);
  }
  void open(const std::string& path) {
    close();
    path_ = path;
    if (has_gz_ext(path)) {
      gz_ = gzopen(path.c_str(), "wb");
      if (gz_) gzbuffer(gz_, 1 << 18);
    } else {
      fp_ = fopen(path.c_str(), "wb");
    }
    if (!gz_ && !fp_) throw Fail("cannot write to file : " + path);
  }
  bool is_open() const { return gz_ || fp_; }
  void write(const char* p, size_t n) {
    if (n == 0) return;
    bool ok;
    if (gz_) {
      ok = true;
      while (n > 0 && ok) {                       // gzwrite takes an unsigned length
        const unsigned c = (unsigned)std::min<size_t>(n, 1u << 30);
        ok = gzwrite(gz_, p, c) == (int)c;
        p += c;
        n -= c;
      }
    } else {
      ok = fwrite(p, 1, n, fp_) == n;
    }
    if (!ok) throw Fail("cannot write to file : " + path_);
  }
  TextWriter& operator<<(const std::string& s) {
    write(s.data(), s.size());
    return *this;
  }
  TextWriter& operator<<(const char* s) {
    write(s, strlen(s));
    return *this;
  }
  void close() {
    bool ok = true;
    if (gz_) { ok = gzclose(gz_) == Z_OK; gz_ = nullptr; }
    if (fp_) { ok = fclose(fp_) == 0; fp_ = nullptr; }
    if (!ok) throw Fail("cannot write to file : " + path_);
  }

 private:
  std::string path_;
  gzFile gz_ = nullptr;
  FILE* fp_ = nullptr;
};

}  // namespace rgh
