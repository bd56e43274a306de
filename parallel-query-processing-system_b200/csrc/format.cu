// format.cu -- K7: SELECT projection rendered on the device
//
// Replaces the per-cell get_attribute_string_value + strdup loop of executeQuerySelectSerial
// (engine/serial/executeEngine-serial.c:216-248, :504-515; 1.9 s for SELECT * over 1 M rows in the
// reference).  For every matching row id and every projected NUMERIC column one thread gathers the value
// and writes its text -- "%llu" / "%d" / "true" | "false", NUL-terminated, exactly what sprintf produces --
// into a FIXED-WIDTH slot (24 / 16 / 8 bytes).  Text columns need no rendering at all: a device text cell
// is NUL-padded to its fixed width, so the plain gather (K2) of the column already is an array of C
// strings.  The host therefore receives, per projected column, one dense array of slots and builds
// data[i][j] = slots_j + i * slot_width_j by arithmetic: no per-cell strlen / memcpy / malloc anywhere.
// Bound: HBM random access for the gather (one 4-128 B cell per id), then PCIe for the slots.

#include "scan_kernels.cuh"

namespace qpe {

namespace {

// decimal digits of v into buf (most significant first); returns the length
__device__ __forceinline__ int render_u64(unsigned long long v, char *buf) {
    char tmp[20];
    int n = 0;
    do {
        const unsigned long long q = v / 10ull;
        tmp[n++] = static_cast<char>('0' + static_cast<int>(v - q * 10ull));
        v = q;
    } while (v);
    for (int i = 0; i < n; ++i) buf[i] = tmp[n - 1 - i];
    return n;
}

__global__ void __launch_bounds__(256) format_u64_kernel(const unsigned long long *__restrict__ col,
                                                         const uint32_t *__restrict__ ids, long long n,
                                                         uint2 *__restrict__ out) {  // 24 B per slot = 3 x uint2
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        alignas(8) char s[24] = {0};
        render_u64(__ldg(col + __ldg(ids + i)), s);  // <= 20 digits: s[20..23] stay NUL
        const uint2 *w = reinterpret_cast<const uint2 *>(s);
        out[3 * i] = w[0];
        out[3 * i + 1] = w[1];
        out[3 * i + 2] = w[2];
    }
}

__global__ void __launch_bounds__(256) format_i32_kernel(const int *__restrict__ col, const uint32_t *__restrict__ ids,
                                                         long long n, uint4 *__restrict__ out) {  // 16 B per slot
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        alignas(16) char s[16] = {0};
        const int v = __ldg(col + __ldg(ids + i));
        int p = 0;
        unsigned int mag = static_cast<unsigned int>(v);
        if (v < 0) {
            s[p++] = '-';
            mag = 0u - mag;  // INT_MIN safe
        }
        render_u64(mag, s + p);  // '-' + 10 digits + NUL = 12 <= 16
        out[i] = *reinterpret_cast<const uint4 *>(s);
    }
}

__global__ void __launch_bounds__(256) format_bool_kernel(const uint8_t *__restrict__ col, const uint32_t *__restrict__ ids,
                                                          long long n, uint2 *__restrict__ out) {  // 8 B per slot
    // little endian: "true\0\0\0\0" / "false\0\0\0"
    const uint2 t = make_uint2(0x65757274u, 0u);           // 't','r','u','e'
    const uint2 f = make_uint2(0x736c6166u, 0x00000065u);  // 'f','a','l','s' | 'e'
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = __ldg(col + __ldg(ids + i)) ? t : f;
}

int grid_of(long long work) {
    long long g = (work + 255) / 256;
    if (g > 148ll * 16) g = 148ll * 16;
    return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace

uint32_t format_slot_width(int col_type, uint32_t cell_width) {
    switch (col_type) {
        case T_U64: return 24;
        case T_I32: return 16;
        case T_BOOL: return 8;
        default: return cell_width;  // text: the NUL-padded cell itself
    }
}

cudaError_t format_launch(const uint8_t *col, int col_type, uint32_t cell_width, const uint32_t *ids, int64_t n,
                          uint8_t *out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    switch (col_type) {
        case T_U64:
            format_u64_kernel<<<grid_of(n), 256, 0, stream>>>(reinterpret_cast<const unsigned long long *>(col), ids, n,
                                                              reinterpret_cast<uint2 *>(out));
            break;
        case T_I32:
            format_i32_kernel<<<grid_of(n), 256, 0, stream>>>(reinterpret_cast<const int *>(col), ids, n,
                                                              reinterpret_cast<uint4 *>(out));
            break;
        case T_BOOL:
            format_bool_kernel<<<grid_of(n), 256, 0, stream>>>(col, ids, n, reinterpret_cast<uint2 *>(out));
            break;
        default:
            return gather_launch(col, cell_width, ids, n, out, stream);
    }
    return cudaGetLastError();
}

}  // namespace qpe
