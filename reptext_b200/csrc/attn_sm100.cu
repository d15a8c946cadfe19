// placeholder, replaced below
#include "rt_internal.h"
namespace rt {
bool attention_tc_supported(const AttnArgs& a, std::string* why) { if (why) *why = "not built yet"; return false; }
void launch_attention_tc(const AttnArgs&, cudaStream_t, int) { throw Error(RT_ERR_UNSUPPORTED, "tcgen05 attention not built"); }
}
