// codec_stub -- LZMA and bzip2 entry points of the reference are stubbed so the vendored lzma/bzip2 trees need
// not be built by the tools that never write them; ref_octree defines REF_REAL_LZMA_BZ2 and links the reference's
// LzmaCompression.cpp / BzlibCompression.cpp with the vendored LZMA SDK and bzip2 sources compiled in place.
// zlib and LZ4 are REAL when REF_REAL_ZLIB_LZ4 is defined:
// the reference's own ZlibCompression.cpp / Lz4Compression.cpp + vendored lz4.c are compiled in place and linked
// with the system libz (ref_octree writes compressed octree files for the reader tests).  Test infrastructure only.
#include <array>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <stdexcept>
static void no_codec() { throw std::runtime_error("oracle/_ref: compression not built"); }
#ifndef REF_REAL_ZLIB_LZ4
void zDecompress(std::shared_ptr<uint8_t>, std::shared_ptr<uint8_t>&, size_t) { no_codec(); }
size_t zCompress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, uint32_t) { no_codec(); return 0; }
#endif
#ifndef REF_REAL_LZMA_BZ2
void lzmaProperties(std::array<uint8_t, 5>&, uint32_t) {}
void lzmaDecompress(std::shared_ptr<uint8_t>, std::shared_ptr<uint8_t>&, size_t, std::array<uint8_t, 5> const&) { no_codec(); }
size_t lzmaCompress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, std::array<uint8_t, 5>&, uint32_t) { no_codec(); return 0; }
#endif
#ifndef REF_REAL_ZLIB_LZ4
void lz4Decompress(std::shared_ptr<uint8_t>, std::shared_ptr<uint8_t>&, size_t) { no_codec(); }
size_t lz4Compress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, uint32_t) { no_codec(); return 0; }
#endif
#ifndef REF_REAL_LZMA_BZ2
void bzDecompress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, size_t) { no_codec(); }
size_t bzCompress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, uint32_t) { no_codec(); return 0; }
#endif
