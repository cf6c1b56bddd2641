// Exercises include/zippy_b200.hpp (the compiled-language host mirror of zippy.nim) end to end
// on a GPU: round trips for every format and level, the seam checksums, and the error contract.
#include <cstdio>
#include <string>

#include "../../include/zippy_b200.hpp"

static uint32_t ref_crc32(const std::string &s) {  // bitwise reference, independent of the library
  uint32_t c = ~0u;
  for (unsigned char b : s) {
    c ^= b;
    for (int i = 0; i < 8; i++) c = (c >> 1) ^ ((0u - (c & 1u)) & 0xedb88320u);
  }
  return ~c;
}

int main() {
  std::string text;
  for (int i = 0; i < 20000; i++) text += "the quick brown fox " + std::to_string(i % 97) + " jumps over the lazy dog\n";
  std::string inputs[] = {std::string(), std::string("a"), text, std::string(300000, 'z')};
  int checked = 0;
  for (const std::string &x : inputs) {
    if (zippy::crc32(x) != ref_crc32(x)) { std::printf("crc mismatch\n"); return 1; }
    for (int level : {-2, -1, 0, 1, 6, 9})
      for (auto fmt : {zippy::dfGzip, zippy::dfZlib, zippy::dfDeflate}) {
        std::string c = zippy::compress(x, level, fmt);
        std::string y = zippy::uncompress(c, fmt == zippy::dfDeflate ? zippy::dfDeflate : zippy::dfDetect);
        if (y != x) { std::printf("round trip failed level %d fmt %d\n", level, (int)fmt); return 1; }
        checked++;
      }
  }
  auto batch = zippy::compressBatch({text, "x", ""}, zippy::BestSpeed, zippy::dfGzip);
  if (zippy::uncompress(batch[0]) != text || zippy::uncompress(batch[1]) != "x" || !zippy::uncompress(batch[2]).empty()) return 1;
  int errors = 0;
  try { zippy::compress(text, 10); } catch (const zippy::ZippyError &) { errors++; }
  try { zippy::uncompress(std::string("definitely not compressed data")); } catch (const zippy::ZippyError &) { errors++; }
  std::string bad = zippy::compress(text, 1, zippy::dfGzip);
  bad[bad.size() - 6] ^= 1;
  try { zippy::uncompress(bad); } catch (const zippy::ZippyError &) { errors++; }
  if (errors != 3) { std::printf("error contract: %d of 3\n", errors); return 1; }
  std::printf("OK %d round trips\n", checked);
  return 0;
}
