// punchout.cuh -- punch-in + punch-out search (FeatureCorrelationImpl.scala:250-393).
#pragma once
#include "common.cuh"
#include "corr.cuh"

namespace sgz {
inline int corr_select_punchout(sgz_corr *, int32_t *) {
  set_error("punch-out search: selection not implemented yet");
  return SGZ_ERR_STATE;
}
inline int corr_merge_punchout(sgz_corr *, const sgz_record *, int32_t, int32_t *) {
  set_error("punch-out search: merge not implemented yet");
  return SGZ_ERR_STATE;
}
}  // namespace sgz
