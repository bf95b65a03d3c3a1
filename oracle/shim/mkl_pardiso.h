/* Forwarding stub: the reference includes <mkl_pardiso.h> (include/lpm.h:18-23); everything lives in mkl.h. Test infrastructure only. */
#include "mkl.h"
