// infidex_b200 -- Stage 2 on device (WordMatcher candidates, lexical coverage, fusion, truncation, filter VM, facets).
#pragma once
#include "ifx_stage1.h"

namespace ifx {

struct FilterProg { const void* consts; const void* code; int32_t n_consts, n_code; };

struct Stage2Buffers { int dummy; };

struct FinalOut { int64_t* key; float* score; uint8_t* tie; int32_t* n; int32_t* total; int32_t* status; int32_t* facet_col; int32_t* facet_val; int32_t* facet_cnt; int32_t* n_facets; int32_t cap, fcap; };

}  // namespace ifx
