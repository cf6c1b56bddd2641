// Exercises include/zippy_b200_tar.hpp: argv[1] = a .tar or .tar.gz made by Python's tarfile; prints one
// line per entry (kind|path|size|crc32|mode|mtime) for the Python test to compare with tarfile's view.
// Linked against libzippy_b200.so on a GPU box, or against mock_abi_zlib.cpp on a CPU-only machine.
#include <cstdio>
#include <fstream>
#include <sstream>

#include "../../include/zippy_b200_tar.hpp"

static uint32_t crc_of(const std::string &s) {
  uint32_t c = ~0u;
  for (unsigned char b : s) {
    c ^= b;
    for (int i = 0; i < 8; i++) c = (c >> 1) ^ ((0u - (c & 1u)) & 0xedb88320u);
  }
  return ~c;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string file = ss.str();
  int errors = 0;
  try { zippy::readTarball(file.substr(0, file.size() / 2)); } catch (const zippy::ZippyError &) { errors++; }   // cut in half
  try { zippy::readTarball(std::string("x")); } catch (const zippy::ZippyError &) { errors++; }
  if (errors != 2) { std::printf("FAILED error contract %d\n", errors); return 1; }
  for (const zippy::TarEntry &e : zippy::readTarball(file))
    std::printf("%c|%s|%zu|%u|%o|%llu\n", e.kind == zippy::TarEntry::File ? 'f' : e.kind == zippy::TarEntry::Directory ? 'd' : 'l',
                e.path.c_str(), e.contents.size(), e.kind == zippy::TarEntry::File ? crc_of(e.contents) : 0u, e.mode & 0777u,
                (unsigned long long)e.mtime);
  std::printf("OK\n");
  return 0;
}
