// Inter-frame mode info, motion vector prediction and variable transform trees (spec 5.11.7 ...).
// Placeholder until the inter path lands: inter frames are reported as unsupported.
#include "../../include/av1r.h"
#include "tile.h"

namespace av1r {

void TileDecoder::inter_frame_mode_info() { fail(AV1R_ENOSYS, "inter frames are not supported yet"); }
void TileDecoder::read_var_tx_size(int, int, int, int) { fail(AV1R_ENOSYS, "inter frames are not supported yet"); }
void TileDecoder::transform_tree(int, int, int, int) { fail(AV1R_ENOSYS, "inter frames are not supported yet"); }

}  // namespace av1r
