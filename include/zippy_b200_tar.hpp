// zippy_b200_tar.hpp -- header-only C++ form of the reference's tarball reader
// (src/zippy/tarballs.nim:25-141) over include/zippy_b200.hpp.
//
// A .tar.gz is ONE gzip member: it goes through zippy::uncompress (GPU path; large members made
// of independent pieces are decoded in parallel), then the 512-byte header walk
// runs on the host as in the reference: ustar prefix, GNU 'L' long names, files / directories /
// symlinks, pax and vendor records skipped, anything else is an error, unsafe paths rejected.
// The Python form (with extraction to disk) is zippy_b200/tarballs.py.
#pragma once
#include <string>
#include <vector>

#include "zippy_b200.hpp"

namespace zippy {

struct TarEntry {
  enum Kind { File, Directory, Symlink } kind = File;
  std::string path;      // prefix/name, or the preceding 'L' record
  std::string contents;  // file bytes, or the link target
  uint32_t mode = 0;
  uint64_t mtime = 0;
};

namespace tardetail {
[[noreturn]] inline void fail(const std::string &msg) { throw ZippyError(ZB200_ERR_UNCOMPRESS, msg); }
// tarballs.nim:5-23: the first run of ASCII digits in the field, base 8 (0 if there is none)
inline uint64_t oct(const std::string &d, size_t pos, size_t n) {
  size_t i = pos, end = pos + n;
  while (i < end && !(d[i] >= '0' && d[i] <= '9')) i++;
  uint64_t v = 0;
  for (; i < end && d[i] >= '0' && d[i] <= '9'; i++) {
    if (d[i] > '7') fail("invalid octal digit");
    v = v * 8 + (uint64_t)(d[i] - '0');
  }
  return v;
}
inline std::string cstr(const std::string &d, size_t pos, size_t n) {
  size_t k = 0;
  while (k < n && d[pos + k] != '\0') k++;
  return d.substr(pos, k);
}
inline void check_safe(const std::string &path) {  // internal.nim verifyPathIsSafeToExtract
  if (!path.empty() && (path[0] == '/' || path[0] == '\\')) fail("Absolute path not allowed " + path);
  if (path.size() > 1 && path[1] == ':') fail("Absolute path not allowed " + path);
  size_t a = 0;
  while (a <= path.size()) {
    size_t b = path.find_first_of("/\\", a);
    if (b == std::string::npos) b = path.size();
    if (path.compare(a, b - a, "..") == 0) fail("Path ../ not allowed " + path);
    a = b + 1;
  }
}
}  // namespace tardetail

// The entries of a .tar or .tar.gz held in memory (tarballs.nim:40-123 without the file system part).
inline std::vector<TarEntry> readTarball(const std::string &file) {
  using namespace tardetail;
  if (file.size() < 2) fail("Invalid buffer, unable to uncompress");
  const bool gz = (unsigned char)file[0] == 31 && (unsigned char)file[1] == 139;
  const std::string data = gz ? uncompress(file, dfGzip) : file;
  std::vector<TarEntry> out;
  std::string longName;
  size_t pos = 0;
  while (pos < data.size()) {
    if (pos + 512 > data.size()) fail("Attempted to read past end of file, corrupted tarball?");
    const std::string name = cstr(data, pos, 100);
    const uint64_t mode = oct(data, pos + 100, 7), size = oct(data, pos + 124, 11), mtime = oct(data, pos + 136, 11);
    const char typeflag = data[pos + 156];
    const std::string linkname = cstr(data, pos + 157, 100);
    const std::string prefix = cstr(data, pos + 257, 6) == "ustar" ? cstr(data, pos + 345, 155) : std::string();
    pos += 512;
    if (pos + size > data.size()) fail("Attempted to read past end of file, corrupted tarball?");
    if (!name.empty() || !longName.empty()) {
      TarEntry e;
      if (!longName.empty()) {
        e.path = longName;
        longName.clear();
      } else {
        e.path = prefix.empty() ? name : prefix + "/" + name;
      }
      check_safe(e.path);
      e.mode = (uint32_t)mode;
      e.mtime = mtime;
      if (typeflag == '0' || typeflag == '\0') {
        e.kind = TarEntry::File;
        e.contents = data.substr(pos, (size_t)size);
        out.push_back(e);
      } else if (typeflag == '5') {
        e.kind = TarEntry::Directory;
        out.push_back(e);
      } else if (typeflag == '2') {
        e.kind = TarEntry::Symlink;
        e.contents = linkname;
        out.push_back(e);
      } else if (typeflag == 'L') {
        longName = cstr(data, pos, (size_t)size);
      } else if (typeflag == 'g' || typeflag == 'x' || (typeflag >= 'A' && typeflag <= 'Z')) {
        // pax and vendor records: skipped, as in the reference
      } else {
        fail(std::string("Unsupported header type ") + typeflag);
      }
    }
    pos += (size_t)((size + 511) & ~(uint64_t)511);
  }
  return out;
}

}  // namespace zippy
