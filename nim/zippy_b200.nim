## zippy_b200.nim -- drop-in replacement for `import zippy` that routes the codec core
## through libzippy_b200.so (include/zippy_b200.h).  Mirrors src/zippy.nim:11-177 of the
## reference: same procs, defaults and ZippyError behaviour; framing stays here on the host
## for the single-input procs, exactly where zippy.nim has it.
##
## NOT COMPILED in this repository's environment (no Nim toolchain on either box); it is
## the binding a maintainer adds.  The ABI it binds is exercised by tests/ through ctypes.
## Complete for the reference's public surface: compress / uncompress (pointer, string and
## seq[uint8] overloads), uncompressGzip, crc32, adler32, plus the batch forms.

import std/sysrand

type
  ZippyError* = object of CatchableError
  CompressedDataFormat* = enum
    dfDetect, dfZlib, dfGzip, dfDeflate
  Zb200Ctx = pointer

const
  NoCompression* = 0
  BestSpeed* = 1
  BestCompression* = 9
  DefaultCompression* = -1
  HuffmanOnly* = -2
  lib = "libzippy_b200.so"

proc zb200_init(device: cint, ctx: ptr Zb200Ctx): cint {.importc, cdecl, dynlib: lib.}
proc zb200_strerror(status: cint): cstring {.importc, cdecl, dynlib: lib.}
proc zb200_deflate_bound(len: csize_t): csize_t {.importc, cdecl, dynlib: lib.}
proc zb200_deflate(ctx: Zb200Ctx, src: pointer, len: csize_t, level: cint,
                   dst: pointer, dstCap: csize_t, dstLen: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc zb200_inflate_size(ctx: Zb200Ctx, src: pointer, len, pos: csize_t,
                        outLen: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc zb200_inflate(ctx: Zb200Ctx, src: pointer, len, pos: csize_t,
                   dst: pointer, dstCap: csize_t, dstLen: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc zb200_decode_begin(ctx: Zb200Ctx, src: pointer, len: csize_t, dataFormat: cint, pos: csize_t,
                        outLen: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc zb200_decode_finish(ctx: Zb200Ctx, dst: pointer, dstCap: csize_t,
                         dstLen: ptr csize_t): cint {.importc, cdecl, dynlib: lib.}
proc zb200_crc32(ctx: Zb200Ctx, src: pointer, len: csize_t, res: ptr uint32): cint {.importc, cdecl, dynlib: lib.}
proc zb200_adler32(ctx: Zb200Ctx, src: pointer, len: csize_t, res: ptr uint32): cint {.importc, cdecl, dynlib: lib.}

var ctx {.threadvar.}: Zb200Ctx

template check(rc: cint) =
  if rc != 0:
    raise newException(ZippyError, $zb200_strerror(rc))

template failUncompress() =
  raise newException(ZippyError, "Invalid buffer, unable to uncompress")   # internal.nim:191-192

proc read32(s: ptr UncheckedArray[uint8], pos: int): uint32 {.inline.} =
  s[pos].uint32 or (s[pos + 1].uint32 shl 8) or (s[pos + 2].uint32 shl 16) or (s[pos + 3].uint32 shl 24)

proc getCtx(): Zb200Ctx =
  if ctx == nil:
    check zb200_init(-1, ctx.addr)
  ctx

proc crc32*(src: pointer, len: int): uint32 =          # crc.nim:53
  check zb200_crc32(getCtx(), src, len.csize_t, result.addr)
proc crc32*(src: string): uint32 = crc32(src.cstring, src.len)
proc adler32*(src: pointer, len: int): uint32 =        # adler32.nim:6
  check zb200_adler32(getCtx(), src, len.csize_t, result.addr)
proc adler32*(src: string): uint32 = adler32(src.cstring, src.len)

proc deflate(dst: var string, src: pointer, len, level: int) =   # deflate.nim:207 (appends)
  let start = dst.len
  dst.setLen(start + zb200_deflate_bound(len.csize_t).int)
  var n: csize_t
  check zb200_deflate(getCtx(), src, len.csize_t, level.cint, dst[start].addr,
                      (dst.len - start).csize_t, n.addr)
  dst.setLen(start + n.int)

proc inflate(dst: var string, src: pointer, len, pos: int) =     # inflate.nim:268
  ## one decode: the library inflates into its own device memory and reports the size
  ## (zb200_decode_begin), then copies the bytes into the string (zb200_decode_finish)
  var n: csize_t
  check zb200_decode_begin(getCtx(), src, len.csize_t, dfDeflate.cint, pos.csize_t, n.addr)
  dst.setLen(n.int)
  var dummy: char
  check zb200_decode_finish(getCtx(), (if n > 0: dst[0].addr else: dummy.addr), n, n.addr)

proc compress*(src: pointer, len: int, level = DefaultCompression,
               dataFormat = dfGzip): string {.raises: [ZippyError].} =
  ## zippy.nim:11-84, framing unchanged
  case dataFormat
  of dfGzip:
    result.setLen(10)
    result[0] = 31.char; result[1] = 139.char; result[2] = 8.char; result[3] = (1 shl 3).char
    var urand: array[1, uint8]
    if not urandom(urand):
      raise newException(ZippyError, "Failed to generate random number")
    for i in 0 ..< (urand[0] mod 26).int: result.add (97 + i).char
    result.add '\0'
    deflate(result, src, len, level)
    let checksum = crc32(src, len)
    for s in [0, 8, 16, 24]: result.add(((checksum shr s) and 255).char)
    for s in [0, 8, 16, 24]: result.add(((len shr s) and 255).char)
  of dfZlib:
    result.setLen(2)
    result[0] = 0x78.char; result[1] = 0x01.char
    deflate(result, src, len, level)
    let checksum = adler32(src, len)
    for s in [24, 16, 8, 0]: result.add(((checksum shr s) and 255).char)
  of dfDeflate:
    deflate(result, src, len, level)
  else:
    raise newException(ZippyError, "Invalid data format " & $dfDetect)

proc compress*(src: string, level = DefaultCompression, dataFormat = dfGzip): string =
  compress(src.cstring, src.len, level, dataFormat)

proc compress*(src: seq[uint8], level = DefaultCompression,
               dataFormat = dfGzip): seq[uint8] {.inline, raises: [ZippyError].} =
  ## zippy.nim:93-98: the seq overload shares the string's buffer
  cast[seq[uint8]](compress(cast[string](src), level, dataFormat))

proc uncompressGzip*(dst: var string, src: pointer, len: int, trustSize = false) {.raises: [ZippyError].} =
  ## gzip.nim:3-88: header checks, inflate, then CRC-32 and ISIZE from the LAST 8 bytes of the buffer
  ## (`trustSize` only pre-sizes `dst` in the reference, gzip.nim:72-76; here the library sizes it).
  if len < 18: failUncompress()
  let src = cast[ptr UncheckedArray[uint8]](src)
  let
    id1 = src[0]; id2 = src[1]; cm = src[2]; flg = src[3]
  if id1 != 31 or id2 != 139:
    raise newException(ZippyError, "Failed gzip identification values check")
  if cm != 8: raise newException(ZippyError, "Unsupported compression method")
  if (flg and 0b11100000) > 0.uint8: raise newException(ZippyError, "Reserved flag bits set")
  let
    fhcrc = (flg and (1.uint8 shl 1)) != 0
    fextra = (flg and (1.uint8 shl 2)) != 0
    fname = (flg and (1.uint8 shl 3)) != 0
    fcomment = (flg and (1.uint8 shl 4)) != 0
  var pos = 10
  if fextra: raise newException(ZippyError, "Currently unsupported flags are set")
  proc nextZeroByte(src: ptr UncheckedArray[uint8], len, start: int): int =
    for i in start ..< len:
      if src[i] == 0: return i
    failUncompress()
  if fname: pos = nextZeroByte(src, len, pos) + 1
  if fcomment: pos = nextZeroByte(src, len, pos) + 1
  if fhcrc:
    if pos + 2 >= len: failUncompress()
    pos += 2                               # not verified (gzip.nim:55-59)
  if pos + 8 >= len: failUncompress()
  let
    checksum = read32(src, len - 8)
    isize = read32(src, len - 4)
  inflate(dst, src, len, pos)
  if checksum != crc32(dst): raise newException(ZippyError, "Checksum verification failed")
  if isize != (dst.len mod (1 shl 32)).uint32: raise newException(ZippyError, "Size verification failed")

proc uncompress*(src: pointer, len: int, dataFormat = dfDetect): string {.raises: [ZippyError].} =
  ## zippy.nim:100-165, framing unchanged: detect, header checks, inflate, trailer verification
  let src = cast[ptr UncheckedArray[uint8]](src)
  case dataFormat
  of dfDetect:
    if len > 18 and src[0] == 31 and src[1] == 139 and src[2] == 8 and (src[3] and 0b11100000) == 0:
      return uncompress(src, len, dfGzip)
    if len > 6 and (src[0] and 0b00001111) == 8 and (src[0] shr 4) <= 7 and
        ((src[0].uint16 * 256) + src[1].uint16) mod 31 == 0:
      return uncompress(src, len, dfZlib)
    raise newException(ZippyError, "Unable to detect compressed data format")
  of dfGzip:
    uncompressGzip(result, src, len)
  of dfZlib:
    if len < 6: failUncompress()
    let
      cmf = src[0]; flg = src[1]
      cm = cmf and 0b00001111
      cinfo = cmf shr 4
    if cm != 8: raise newException(ZippyError, "Unsupported compression method")
    if cinfo > 7.uint8: raise newException(ZippyError, "Invalid compression info")
    if ((cmf.uint16 * 256) + flg.uint16) mod 31 != 0: raise newException(ZippyError, "Invalid header")
    if (flg and 0b00100000) != 0: raise newException(ZippyError, "Preset dictionary is not yet supported")
    inflate(result, src, len, 2)
    let checksum = (src[len - 4].uint32 shl 24) or (src[len - 3].uint32 shl 16) or
                   (src[len - 2].uint32 shl 8) or src[len - 1].uint32
    if checksum != adler32(result): raise newException(ZippyError, "Checksum verification failed")
  of dfDeflate:
    inflate(result, src, len, 0)

proc uncompress*(src: string, dataFormat = dfDetect): string {.inline, raises: [ZippyError].} =
  uncompress(src.cstring, src.len, dataFormat)          # zippy.nim:167-171

proc uncompress*(src: seq[uint8], dataFormat = dfDetect): seq[uint8] {.inline, raises: [ZippyError].} =
  cast[seq[uint8]](uncompress(cast[string](src), dataFormat))   # zippy.nim:173-177

# ---- batch (no counterpart in zippy.nim; what ziparchives.nim:505-540 should call instead of a
# per-entry loop of crc32 + compress): N independent inputs, one GPU launch sequence ----
proc zb200_compress_bound(len: csize_t, dataFormat: cint): csize_t {.importc, cdecl, dynlib: lib.}
proc zb200_compress_batch(ctx: Zb200Ctx, srcBase: pointer, srcOffsets: ptr uint64, n: csize_t,
                          level, dataFormat: cint, fnameLens: pointer,
                          dstBase: pointer, dstCap: csize_t, dstOffsets: ptr uint64,
                          statuses: ptr cint): cint {.importc, cdecl, dynlib: lib.}

proc compressBatch*(items: openArray[string], level = DefaultCompression,
                    dataFormat = dfGzip): seq[string] {.raises: [ZippyError].} =
  var
    base: string
    offs = newSeq[uint64](items.len + 1)
    outOffs = newSeq[uint64](items.len + 1)
    bound = 64
  for i, item in items:
    base.add item
    offs[i + 1] = base.len.uint64
    bound += zb200_compress_bound(item.len.csize_t, dataFormat.cint).int + 64
  var dst = newString(bound)
  if base.len == 0: base.add '\0'
  check zb200_compress_batch(getCtx(), base[0].addr, offs[0].addr, items.len.csize_t,
                             level.cint, dataFormat.cint, nil, dst[0].addr, dst.len.csize_t,
                             outOffs[0].addr, nil)
  for i in 0 ..< items.len:
    result.add dst[outOffs[i].int ..< outOffs[i + 1].int]
