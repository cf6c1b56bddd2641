## zippy_b200.nim -- drop-in replacement for `import zippy` that routes the codec core
## through libzippy_b200.so (include/zippy_b200.h).  Mirrors src/zippy.nim:11-177 of the
## reference: same procs, defaults and ZippyError behaviour; framing stays here on the host
## for the single-input procs, exactly where zippy.nim has it.
##
## NOT COMPILED in this repository's environment (no Nim toolchain on either box); it is
## the binding a maintainer adds.  The ABI it binds is exercised by tests/ through ctypes.

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
proc zb200_crc32(ctx: Zb200Ctx, src: pointer, len: csize_t, res: ptr uint32): cint {.importc, cdecl, dynlib: lib.}
proc zb200_adler32(ctx: Zb200Ctx, src: pointer, len: csize_t, res: ptr uint32): cint {.importc, cdecl, dynlib: lib.}

var ctx {.threadvar.}: Zb200Ctx

template check(rc: cint) =
  if rc != 0:
    raise newException(ZippyError, $zb200_strerror(rc))

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
  var n: csize_t
  check zb200_inflate_size(getCtx(), src, len.csize_t, pos.csize_t, n.addr)
  dst.setLen(n.int)
  if n > 0:
    check zb200_inflate(getCtx(), src, len.csize_t, pos.csize_t, dst[0].addr, n, n.addr)

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

# uncompress*: zippy.nim:100-165 and gzip.nim:3-88 carry over verbatim with `inflate`,
# `crc32`, `adler32` bound as above (header checks and trailer verification stay host-side).

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
