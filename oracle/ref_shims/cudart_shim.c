/* TEST INFRASTRUCTURE.  Three CUDA-runtime entry points, faked, so that the
 * reference's CPU-side weight pre-processor (cutlass_preprocessors.cc:115-131
 * calls getSMVersion(), cuda_utils.h:281-290) can be linked and run on a host
 * with no GPU -- and on the B200 box, where the real answer (sm_100) makes the
 * reference throw "Unsupported Arch".  The shim answers "sm_80", the layout the
 * reference's published numbers (A100) and its *.q.bin files use.            */
#include <stddef.h>
int cudaGetDevice(int* device) { *device = 0; return 0; }
int cudaDeviceGetAttribute(int* value, int attr, int device)
{
    (void)device;
    /* cudaDevAttrComputeCapabilityMajor = 75, Minor = 76 */
    if (attr == 75) { *value = 8; } else if (attr == 76) { *value = 0; } else { *value = 0; }
    return 0;
}
const char* cudaGetErrorString(int err) { (void)err; return "cudart_shim"; }
