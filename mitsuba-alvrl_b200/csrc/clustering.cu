/* clustering.cu -- placeholder, replaced below */
#include "context.h"
namespace alvrl {
void build_clusters_device(alvrl_ctx *, bool) { throw Error(ALVRL_ERR_UNSUPPORTED, "device clustering not built yet"); }
}
