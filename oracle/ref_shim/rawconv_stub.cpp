// rawconv_stub -- AbstrConverter's quantisation helpers reference RAWConverter::ConvertRAWDataset (the whole import
// pipeline, which needs the IO manager); oracle/_ref/ref_dataset only READS .uvf files and never reaches it.
#include <stdexcept>
#include "IO/RAWConverter.h"
bool RAWConverter::ConvertRAWDataset(const std::string&, const std::string&, const std::string&, uint64_t, unsigned,
                                     uint64_t, uint64_t, bool, bool, bool, UINT64VECTOR3, FLOATVECTOR3, const std::string&,
                                     const std::string&, const uint64_t, const uint64_t, const bool, const bool, uint32_t,
                                     uint32_t, uint32_t, KVPairs*, const bool) {
  throw std::runtime_error("oracle/_ref: RAWConverter is not built");
}
