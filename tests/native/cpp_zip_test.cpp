// Exercises include/zippy_b200_zip.hpp: argv[1] = an archive made by Python's zipfile (stored and
// deflated entries, a directory, a comment, junk in front), argv[2] = where to write an archive made
// here for zipfile to check.  Linked against libzippy_b200.so on a GPU box, or against
// mock_abi_zlib.cpp on a CPU-only machine (container logic only).
#include <cstdio>
#include <fstream>
#include <sstream>

#include "../../include/zippy_b200_zip.hpp"

static std::string slurp(const char *path) {
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
#define REQUIRE(c)                                             \
  do {                                                         \
    if (!(c)) {                                                \
      std::printf("FAILED line %d: %s\n", __LINE__, #c);       \
      return 1;                                                \
    }                                                          \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  // ---- read an archive written by another implementation ----
  zippy::ZipArchiveReader rd(slurp(argv[1]));
  auto names = rd.walkFiles();
  REQUIRE(names.size() == 3 && names[0] == "a/b.txt" && names[1] == "stored.bin" && names[2] == "caf\xc3\xa9.txt");
  auto files = rd.extractFiles(names);
  std::string want;
  for (int i = 0; i < 1000; i++) want += "hello ";
  REQUIRE(files["a/b.txt"] == want);
  REQUIRE(files["stored.bin"] == std::string("\x00\x01\x02", 3));
  REQUIRE(rd.extractFile("caf\xc3\xa9.txt") == "na\xc3\xafve");
  int errors = 0;
  try { rd.extractFile("a/"); } catch (const zippy::ZippyError &) { errors++; }        // a directory
  try { rd.extractFile("missing"); } catch (const zippy::ZippyError &) { errors++; }
  try { zippy::ZipArchiveReader bad(std::string("not a zip archive at all, not even close")); } catch (const zippy::ZippyError &) { errors++; }
  // ---- write one, read it back ----
  std::vector<std::pair<std::string, std::string>> entries = {
      {"README.txt", "Hello, World!"}, {"dir/empty.bin", ""}, {"dir/sub/data.bin", std::string()}, {"caf\xc3\xa9.txt", "na\xc3\xafve"}};
  for (int i = 0; i < 300 * 256; i++) entries[2].second.push_back((char)(i & 255));
  const std::string blob = zippy::createZipArchive(entries);
  zippy::ZipArchiveReader back(blob);
  auto order = back.walkFiles();
  REQUIRE(order.size() == 4 && order[0] == entries[3].first && order[3] == entries[0].first);   // last key first
  auto got = back.extractFiles(order);
  for (auto &e : entries) REQUIRE(got[e.first] == e.second);
  std::string corrupt = blob;
  corrupt[back.records()[1].headerOffset + 30 + entries[2].first.size() + 20 + 9] ^= 0x55;     // payload of data.bin
  try { zippy::ZipArchiveReader(corrupt).extractFile(entries[2].first); } catch (const zippy::ZippyError &) { errors++; }
  try { zippy::createZipArchive({{"", "x"}}); } catch (const zippy::ZippyError &) { errors++; }
  try { zippy::createZipArchive({{"/abs", "x"}}); } catch (const zippy::ZippyError &) { errors++; }
  REQUIRE(errors == 6);
  std::ofstream(argv[2], std::ios::binary) << blob;
  std::printf("OK %zu entries read, %zu bytes written\n", names.size(), blob.size());
  return 0;
}
