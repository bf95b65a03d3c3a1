/* Forwarding stub: the reference includes <mkl_spblas.h> (include/lpm.h:18-23); everything lives in mkl.h. Test infrastructure only. */
#include "mkl.h"
