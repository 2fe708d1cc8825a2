// codec_stub -- the oracle/_ref tools only ever write and read UNCOMPRESSED
// bricks (CT_NONE); the reference's compression entry points are stubbed so the
// vendored zlib/lzma/lz4/bzip2 trees need not be built.  Test infrastructure only.
#include <array>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <stdexcept>
static void no_codec() { throw std::runtime_error("oracle/_ref: compression not built"); }
void zDecompress(std::shared_ptr<uint8_t>, std::shared_ptr<uint8_t>&, size_t) { no_codec(); }
size_t zCompress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, uint32_t) { no_codec(); return 0; }
void lzmaProperties(std::array<uint8_t, 5>&, uint32_t) {}
void lzmaDecompress(std::shared_ptr<uint8_t>, std::shared_ptr<uint8_t>&, size_t, std::array<uint8_t, 5> const&) { no_codec(); }
size_t lzmaCompress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, std::array<uint8_t, 5>&, uint32_t) { no_codec(); return 0; }
void lz4Decompress(std::shared_ptr<uint8_t>, std::shared_ptr<uint8_t>&, size_t) { no_codec(); }
size_t lz4Compress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, uint32_t) { no_codec(); return 0; }
void bzDecompress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, size_t) { no_codec(); }
size_t bzCompress(std::shared_ptr<uint8_t>, size_t, std::shared_ptr<uint8_t>&, uint32_t) { no_codec(); return 0; }
