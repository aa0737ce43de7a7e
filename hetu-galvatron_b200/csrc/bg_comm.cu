// bg_comm.cu -- symmetric arena, groups, device-side barriers and the peer-memory collectives
// (SURVEY 2.3 rows C1-C3, C5-C14, C16).  sm_100a; NVLink 5 / NVSwitch peer loads & stores, no NCCL.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include <cuda.h>

#include "bg_common.cuh"

namespace bg {
thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};
Tunables g_tun;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
}  // namespace bg

using namespace bg;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct Group {
    int n = 0, me = -1, slot = -1;
    int ranks[BG_MAX_PEERS];
};

struct bg_ctx {
    int rank = 0, world = 1, device = 0;
    char* arena = nullptr;
    size_t arena_bytes = 0, bump = 0, pad_bytes = 0;
    char* peer_base[BG_MAX_WORLD];
    bool peer_ipc[BG_MAX_WORLD];
    std::vector<Group> groups;
    std::map<std::vector<int>, int> gid_of;
    std::map<std::tuple<int, int, int>, int> slot_of;  // (first, stride, size) -> signal slot
    int* err_host = nullptr;                            // mapped pinned: device-side timeout report
    int* err_dev = nullptr;
    unsigned long long p2p_sent[BG_MAX_WORLD][64] = {};
    std::mutex mu;
    // ---- VMM arena / NVLS multicast (opt-in) ----
    bool vmm = false;
    int mc_supported = 0;
    size_t vmm_gran = 0, mc_gran = 0;
    CUmemGenericAllocationHandle arena_handle = 0;
    CUmemGenericAllocationHandle peer_handle[BG_MAX_WORLD] = {};
    struct McGroup {
        CUmemGenericAllocationHandle mc = 0;
        CUdeviceptr va = 0;
        size_t bytes = 0, arena_off = 0;
        bool bound = false;
    };
    std::map<int, McGroup> mc_of;   // gid -> multicast object over the group's NVLS buffer
};

// signal pad: pad[slot][lane][channel][BG_MAX_PEERS] u32, followed by the p2p flags [BG_MAX_WORLD][P2P_FLAGS]
static constexpr size_t kSlotBytes = (size_t)BG_LANES * BG_MAX_CHANNELS * BG_MAX_PEERS * sizeof(uint32_t);
static constexpr int kP2PFlags = 64;

static void enumerate_slots(bg_ctx* c) {
    // every arithmetic progression of 2..BG_MAX_PEERS ranks inside [0, world): same table on all ranks
    int next = 0;
    for (int stride = 1; stride < c->world; ++stride)
        for (int size = 2; size <= BG_MAX_PEERS; ++size)
            for (int first = 0; first + (size - 1) * stride < c->world; ++first)
                c->slot_of[std::make_tuple(first, stride, size)] = next++;
}


// =================================================================================================================
// VMM arena + NVLS (NVSwitch multicast) all-reduce.  OPT-IN (bg_ctx_create_ex flag BG_CTX_VMM, HGB_NVLS=1 on the host side):
// the default arena is cudaMalloc + cudaIpc.  Multicast objects can only bind memory that was created with cuMemCreate, so
// this mode allocates the arena through the virtual-memory API and shares it (and the multicast objects) between the
// processes as POSIX file descriptors, which the host passes over a unix socket (SCM_RIGHTS).
// Replaces NCCL's NVLS all-reduce for the tensor-parallel reductions (mappings_group.py:19, layers.py:474-480):
// two-shot -- every member reduces ITS slice in the switch (multimem.ld_reduce) and broadcasts it (multimem.st).
// =================================================================================================================

namespace {

struct Drv {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    bool ok = false;
};

template <typename F>
bool drv_sym(const char* name, F* out) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *out = reinterpret_cast<F>(p);
    return true;
}

Drv& drv() {
    static Drv d;
    static std::once_flag once;
    std::call_once(once, [] {
        bool ok = true;
        ok &= drv_sym("cuMemCreate", &d.MemCreate);
        ok &= drv_sym("cuMemRelease", &d.MemRelease);
        ok &= drv_sym("cuMemAddressReserve", &d.MemAddressReserve);
        ok &= drv_sym("cuMemAddressFree", &d.MemAddressFree);
        ok &= drv_sym("cuMemMap", &d.MemMap);
        ok &= drv_sym("cuMemUnmap", &d.MemUnmap);
        ok &= drv_sym("cuMemSetAccess", &d.MemSetAccess);
        ok &= drv_sym("cuMemExportToShareableHandle", &d.MemExportToShareableHandle);
        ok &= drv_sym("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle);
        ok &= drv_sym("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity);
        ok &= drv_sym("cuMulticastCreate", &d.MulticastCreate);
        ok &= drv_sym("cuMulticastAddDevice", &d.MulticastAddDevice);
        ok &= drv_sym("cuMulticastBindMem", &d.MulticastBindMem);
        ok &= drv_sym("cuMulticastGetGranularity", &d.MulticastGetGranularity);
        ok &= drv_sym("cuDeviceGet", &d.DeviceGet);
        ok &= drv_sym("cuDeviceGetAttribute", &d.DeviceGetAttribute);
        ok &= drv_sym("cuGetErrorString", &d.GetErrorString);
        d.ok = ok;
    });
    return d;
}

int drv_fail(const char* what, CUresult r) {
    const char* msg = nullptr;
    if (drv().GetErrorString) drv().GetErrorString(r, &msg);
    return fail(BG_ECUDA, "%s: %s (CUresult %d)", what, msg ? msg : "?", (int)r);
}
#define BG_DRV(expr)                                   \
    do {                                               \
        CUresult _r = (expr);                          \
        if (_r != CUDA_SUCCESS) return drv_fail(#expr, _r); \
    } while (0)

CUmemAllocationProp arena_prop(int device) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

int map_rw(CUdeviceptr* va, size_t bytes, size_t align, CUmemGenericAllocationHandle h, int device) {
    Drv& d = drv();
    BG_DRV(d.MemAddressReserve(va, bytes, align, 0, 0));
    BG_DRV(d.MemMap(*va, bytes, 0, h, 0));
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    BG_DRV(d.MemSetAccess(*va, bytes, &acc, 1));
    return BG_OK;
}

}  // namespace

extern "C" int bg_abi_version(void) { return BG_ABI_VERSION; }
extern "C" const char* bg_last_error(void) { return g_last_error.c_str(); }
extern "C" unsigned long long bg_launch_count(void) { return g_launches.load(); }

static long long* tunable(const char* name) {
    if (!name) return nullptr;
    if (!strcmp(name, "comm_ctas")) return &g_tun.comm_ctas;
    if (!strcmp(name, "local_ctas")) return &g_tun.local_ctas;
    if (!strcmp(name, "timeout_ms")) return &g_tun.timeout_ms;
    if (!strcmp(name, "oneshot_bytes")) return &g_tun.oneshot_bytes;
    return nullptr;
}
extern "C" int bg_set_tunable(const char* name, long long value) {
    long long* t = tunable(name);
    if (!t) return fail(BG_EINVAL, "unknown tunable %s", name ? name : "(null)");
    if (t == &g_tun.comm_ctas && (value < 1 || value > BG_MAX_CHANNELS))
        return fail(BG_EINVAL, "comm_ctas must be in [1,%d]", BG_MAX_CHANNELS);
    if (value < 1) return fail(BG_EINVAL, "tunable %s must be positive", name);
    *t = value;
    return BG_OK;
}
extern "C" long long bg_get_tunable(const char* name) {
    long long* t = tunable(name);
    return t ? *t : -1;
}

extern "C" int bg_ctx_create_ex(int rank, int world, int device, size_t arena_bytes, unsigned flags, bg_ctx_t* out) {
    if (!out || world < 1 || world > BG_MAX_WORLD || rank < 0 || rank >= world)
        return fail(BG_EINVAL, "bg_ctx_create: bad rank/world %d/%d", rank, world);
    BG_CUDA(cudaSetDevice(device));
    BG_CUDA(cudaFree(0));   // the primary context exists before any driver-API call
    bg_ctx* c = new bg_ctx();
    c->rank = rank; c->world = world; c->device = device;
    memset(c->peer_base, 0, sizeof(c->peer_base));
    memset(c->peer_ipc, 0, sizeof(c->peer_ipc));
    enumerate_slots(c);
    size_t pad = c->slot_of.size() * kSlotBytes + (size_t)BG_MAX_WORLD * kP2PFlags * 2 * sizeof(uint32_t);
    pad = (pad + 4095) / 4096 * 4096;
    c->pad_bytes = pad;
    c->arena_bytes = pad + ((arena_bytes + 4095) / 4096 * 4096);
    if (flags & BG_CTX_VMM) {
        Drv& d = drv();
        int rc = BG_OK;
        if (!d.ok) rc = fail(BG_ECUDA, "the driver does not expose the virtual-memory / multicast entry points");
        CUmemAllocationProp prop = arena_prop(device);
        size_t gran = 0;
        CUdevice dev = 0;
        if (!rc && d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS)
            rc = fail(BG_ECUDA, "cuMemGetAllocationGranularity failed");
        if (!rc && d.DeviceGet(&dev, device) == CUDA_SUCCESS)
            d.DeviceGetAttribute(&c->mc_supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
        if (!rc && c->mc_supported) {
            CUmulticastObjectProp mp = {};
            mp.numDevices = world > 1 ? (unsigned)world : 2u;
            mp.size = gran;
            mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
            size_t mg = 0;
            if (d.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg) {
                c->mc_gran = mg;
                if (mg > gran) gran = mg;
            } else {
                c->mc_supported = 0;
            }
        }
        if (!rc) {
            c->vmm = true;
            c->vmm_gran = gran;
            c->arena_bytes = (c->arena_bytes + gran - 1) / gran * gran;
            CUresult r = d.MemCreate(&c->arena_handle, c->arena_bytes, &prop, 0);
            if (r != CUDA_SUCCESS) rc = drv_fail("cuMemCreate(arena)", r);
        }
        CUdeviceptr va = 0;
        if (!rc) rc = map_rw(&va, c->arena_bytes, c->vmm_gran, c->arena_handle, device);
        if (rc) { delete c; return rc; }
        c->arena = (char*)va;
    } else {
        cudaError_t e = cudaMalloc(&c->arena, c->arena_bytes);
        if (e != cudaSuccess) {
            size_t want = c->arena_bytes;
            delete c;
            return fail(BG_ENOMEM, "arena cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
        }
    }
    BG_CUDA(cudaMemset(c->arena, 0, pad));
    BG_CUDA(cudaHostAlloc(&c->err_host, 8 * sizeof(int), cudaHostAllocMapped));   // [0] status, [1..7] who/where
    for (int i = 0; i < 8; ++i) c->err_host[i] = 0;
    BG_CUDA(cudaHostGetDevicePointer(&c->err_dev, c->err_host, 0));
    BG_CUDA(cudaDeviceSynchronize());
    c->bump = pad;
    c->peer_base[rank] = c->arena;
    *out = c;
    return BG_OK;
}

extern "C" int bg_ctx_create(int rank, int world, int device, size_t arena_bytes, bg_ctx_t* out) {
    return bg_ctx_create_ex(rank, world, device, arena_bytes, 0u, out);
}

extern "C" int bg_ctx_destroy(bg_ctx_t c) {
    if (!c) return BG_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->vmm) {
        Drv& d = drv();
        for (auto& kv : c->mc_of) {
            if (kv.second.va) { d.MemUnmap(kv.second.va, kv.second.bytes); d.MemAddressFree(kv.second.va, kv.second.bytes); }
            if (kv.second.mc) d.MemRelease(kv.second.mc);
        }
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank && c->peer_handle[r]) {
                d.MemUnmap((CUdeviceptr)c->peer_base[r], c->arena_bytes);
                d.MemAddressFree((CUdeviceptr)c->peer_base[r], c->arena_bytes);
                d.MemRelease(c->peer_handle[r]);
            }
        if (c->arena) { d.MemUnmap((CUdeviceptr)c->arena, c->arena_bytes); d.MemAddressFree((CUdeviceptr)c->arena, c->arena_bytes); }
        if (c->arena_handle) d.MemRelease(c->arena_handle);
    } else {
        for (int r = 0; r < c->world; ++r)
            if (c->peer_ipc[r] && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
        if (c->arena) cudaFree(c->arena);
    }
    if (c->err_host) cudaFreeHost(c->err_host);
    delete c;
    return BG_OK;
}

extern "C" int bg_arena_info(bg_ctx_t c, void** base, size_t* bytes, size_t* used) {
    if (!c) return fail(BG_EINVAL, "null ctx");
    if (base) *base = c->arena;
    if (bytes) *bytes = c->arena_bytes;
    if (used) *used = c->bump;
    return BG_OK;
}

extern "C" int bg_arena_alloc(bg_ctx_t c, size_t bytes, size_t* offset) {
    if (!c || !offset) return fail(BG_EINVAL, "null arg");
    std::lock_guard<std::mutex> lk(c->mu);
    size_t off = (c->bump + 255) / 256 * 256;
    if (off + bytes > c->arena_bytes)
        return fail(BG_ENOMEM, "arena exhausted: want %zu at %zu of %zu (raise arena_bytes)", bytes, off, c->arena_bytes);
    c->bump = off + bytes;
    *offset = off;
    return BG_OK;
}

extern "C" int bg_arena_alloc_aligned(bg_ctx_t c, size_t bytes, size_t align, size_t* offset) {
    if (!c || !offset || !align) return fail(BG_EINVAL, "null arg");
    std::lock_guard<std::mutex> lk(c->mu);
    size_t off = (c->bump + align - 1) / align * align;
    if (off + bytes > c->arena_bytes)
        return fail(BG_ENOMEM, "arena exhausted: want %zu at %zu of %zu (raise arena_bytes)", bytes, off, c->arena_bytes);
    c->bump = off + bytes;
    *offset = off;
    return BG_OK;
}

extern "C" int bg_arena_mode(bg_ctx_t c, int* vmm, int* multicast, size_t* mc_granularity) {
    if (!c) return fail(BG_EINVAL, "null ctx");
    if (vmm) *vmm = c->vmm ? 1 : 0;
    if (multicast) *multicast = c->vmm ? c->mc_supported : 0;
    if (mc_granularity) *mc_granularity = c->mc_gran;
    return BG_OK;
}

extern "C" int bg_arena_export_fd(bg_ctx_t c, int* fd) {
    if (!c || !fd) return fail(BG_EINVAL, "null arg");
    if (!c->vmm) return fail(BG_EINVAL, "bg_arena_export_fd needs a BG_CTX_VMM context (cudaMalloc arenas use bg_arena_export)");
    BG_CUDA(cudaSetDevice(c->device));
    BG_DRV(drv().MemExportToShareableHandle(fd, c->arena_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    return BG_OK;
}

extern "C" int bg_arena_import_fd(bg_ctx_t c, int peer, int fd) {
    if (!c || peer < 0 || peer >= c->world || fd < 0) return fail(BG_EINVAL, "bad peer %d / fd %d", peer, fd);
    if (!c->vmm) return fail(BG_EINVAL, "bg_arena_import_fd needs a BG_CTX_VMM context");
    if (peer == c->rank) return BG_OK;
    BG_CUDA(cudaSetDevice(c->device));
    Drv& d = drv();
    BG_DRV(d.MemImportFromShareableHandle(&c->peer_handle[peer], (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr va = 0;
    int rc = map_rw(&va, c->arena_bytes, c->vmm_gran, c->peer_handle[peer], c->device);   // every rank's arena has the same size
    if (rc) return rc;
    c->peer_base[peer] = (char*)va;
    return BG_OK;
}

extern "C" int bg_arena_export(bg_ctx_t c, void* handle64) {
    if (!c || !handle64) return fail(BG_EINVAL, "null arg");
    if (c->vmm) return fail(BG_EINVAL, "a BG_CTX_VMM arena is shared with bg_arena_export_fd");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    BG_CUDA(cudaSetDevice(c->device));
    BG_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, c->arena));
    return BG_OK;
}

extern "C" int bg_arena_import(bg_ctx_t c, int peer, const void* handle64) {
    if (!c || !handle64 || peer < 0 || peer >= c->world) return fail(BG_EINVAL, "bad peer %d", peer);
    if (peer == c->rank) return BG_OK;
    BG_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    BG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_base[peer] = (char*)p;
    c->peer_ipc[peer] = true;
    return BG_OK;
}

extern "C" int bg_arena_attach_local(bg_ctx_t c, int peer, bg_ctx_t other) {
    if (!c || !other || peer < 0 || peer >= c->world) return fail(BG_EINVAL, "bad peer %d", peer);
    if (other->device != c->device) {
        int can = 0;
        BG_CUDA(cudaDeviceCanAccessPeer(&can, c->device, other->device));
        if (!can) return fail(BG_ENOTMAPPED, "device %d cannot access %d", c->device, other->device);
        BG_CUDA(cudaSetDevice(c->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(BG_ECUDA, "enable peer access: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    c->peer_base[peer] = other->arena;
    return BG_OK;
}

extern "C" int bg_ctx_error_flag(bg_ctx_t c, int* flag) {
    if (!c || !flag) return fail(BG_EINVAL, "null arg");
    *flag = *c->err_host;
    return BG_OK;
}

extern "C" int bg_ctx_error_info(bg_ctx_t c, int* info8) {
    if (!c || !info8) return fail(BG_EINVAL, "null arg");
    for (int i = 0; i < 8; ++i) info8[i] = c->err_host[i];   // mapped host memory: readable after a device trap
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// groups
// ------------------------------------------------------------------------------------------------
extern "C" int bg_group_create(bg_ctx_t c, const int* ranks, int n, int* gid) {
    if (!c || !ranks || !gid || n < 1) return fail(BG_EINVAL, "bg_group_create: bad args");
    if (n > BG_MAX_PEERS) return fail(BG_EGROUP, "group of %d ranks exceeds one NVSwitch domain (%d)", n, BG_MAX_PEERS);
    std::vector<int> key(ranks, ranks + n);
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->gid_of.find(key);
    if (it != c->gid_of.end()) { *gid = it->second; return BG_OK; }
    Group g;
    g.n = n;
    for (int i = 0; i < n; ++i) {
        if (ranks[i] < 0 || ranks[i] >= c->world) return fail(BG_EGROUP, "rank %d outside world %d", ranks[i], c->world);
        if (i && ranks[i] <= ranks[i - 1]) return fail(BG_EGROUP, "rank list must be strictly increasing");
        g.ranks[i] = ranks[i];
        if (ranks[i] == c->rank) g.me = i;
    }
    if (g.me < 0) return fail(BG_EGROUP, "calling rank %d is not a member", c->rank);
    if (n >= 2) {
        int stride = ranks[1] - ranks[0];
        for (int i = 2; i < n; ++i)
            if (ranks[i] - ranks[i - 1] != stride) return fail(BG_EGROUP, "rank list is not an arithmetic progression");
        auto s = c->slot_of.find(std::make_tuple(ranks[0], stride, n));
        if (s == c->slot_of.end()) return fail(BG_EGROUP, "no signal slot for group");
        g.slot = s->second;
    }
    c->groups.push_back(g);
    *gid = (int)c->groups.size() - 1;
    c->gid_of[key] = *gid;
    return BG_OK;
}

extern "C" int bg_group_info(bg_ctx_t c, int gid, int* n, int* my_index, int* ranks_out) {
    if (!c || gid < 0 || gid >= (int)c->groups.size()) return fail(BG_EGROUP, "bad gid %d", gid);
    const Group& g = c->groups[gid];
    if (n) *n = g.n;
    if (my_index) *my_index = g.me;
    if (ranks_out) for (int i = 0; i < g.n; ++i) ranks_out[i] = g.ranks[i];
    return BG_OK;
}

// C mirror of the closed-form membership rules (comm_groups.py:71-236,382-409) for the bit-exact check
static int put_range(int* counts, int* ranks, int first, int stop, int step) {
    int k = 0;
    for (int r = first; r < stop; r += step) ranks[k++] = r;
    *counts = k;
    return k;
}
extern "C" int bg_build_groups(int rank, int world, int pp, int n_layers, const int* tp, const int* sp, const int* cp,
                               int* out_counts, int* out_ranks, int* pp_count, int* pp_ranks) {
    if (world < 1 || world > BG_MAX_WORLD || pp < 1 || world % pp || rank < 0 || rank >= world)
        return fail(BG_EINVAL, "bg_build_groups: bad world/pp/rank");
    const int per_stage = world / pp, base = rank / per_stage * per_stage, local = rank - base;
    for (int i = 0; i < n_layers; ++i) {
        const int t = tp[i], s = sp[i], c = cp[i], mul = t * s;
        if (t < 1 || s < 1 || c < 1 || (t != 1 && s != 1) || per_stage % (mul * c))
            return fail(BG_EINVAL, "layer %d: invalid tp/sp/cp %d/%d/%d", i, t, s, c);
        auto cnt = [&](int kind) { return out_counts + kind * n_layers + i; };
        auto rk = [&](int kind) { return out_ranks + ((size_t)kind * n_layers + i) * BG_MAX_WORLD; };
        put_range(cnt(0), rk(0), rank / t * t, rank / t * t + t, 1);                                   // tp
        put_range(cnt(1), rk(1), rank / s * s, rank / s * s + s, 1);                                   // sp
        int first = base + local / (mul * c) * (mul * c) + local % mul;
        put_range(cnt(2), rk(2), first, first + mul * c, mul);                                         // cp
        put_range(cnt(3), rk(3), base + local % (mul * c), base + per_stage, mul * c);                 // dp
        if (t == 1) put_range(cnt(4), rk(4), base, base + per_stage, 1);                               // sdp
        else put_range(cnt(4), rk(4), base + local % t, base + per_stage, t);
    }
    put_range(pp_count, pp_ranks, rank % per_stage, world, per_stage);
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// device-side cross-rank barrier (per CTA channel): CAS put 0->1 on the peer, CAS wait 1->0 locally
// ------------------------------------------------------------------------------------------------
struct Sig {
    uint32_t* local;               // my pad for (slot, lane): [channel][BG_MAX_PEERS]
    uint32_t* peer[BG_MAX_PEERS];  // the same region in every member's arena
    int me, n;
    unsigned long long timeout_ns;
    int* err;
};

struct PeerPtrs {
    char* p[BG_MAX_PEERS];
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void sig_spin_cas(uint32_t* addr, uint32_t expect, uint32_t desired, bool release,
                                             const Sig& s) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (true) {
        uint32_t old;
        if (release)
            asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
        else
            asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
        if (old == expect) return;
        if ((++spins & 0xff) == 0) {
            unsigned long long now = gtimer();
            if (t0 == 0) t0 = now;
            else if (now - t0 > s.timeout_ns) {
                // who/where: kind 1 = signal a peer (its flag never drained), 2 = wait for a peer's signal
                if (atomicCAS(s.err + 1, 0, release ? 1 : 2) == 0) {
                    s.err[2] = (int)blockIdx.x; s.err[3] = (int)threadIdx.x; s.err[4] = (int)old; s.err[5] = s.me; s.err[6] = s.n;
                }
                *s.err = BG_ETIMEOUT;
                __threadfence_system();
                __trap();
            }
        }
    }
}

// All threads of the CTA call this.
//   kSyncBefore: the whole CTA must have finished its prior loads/stores before the signal is raised
//   kFence:      this CTA wrote data that peers read after the barrier (make it visible at .sys scope)
//   kSyncAfter:  the whole CTA must wait for the barrier before continuing
template <bool kSyncBefore, bool kFence, bool kSyncAfter>
__device__ __forceinline__ void sync_peers(const Sig& s) {
    if (s.n == 1) return;
    if (kFence) __threadfence_system();
    if (kSyncBefore) __syncthreads();
    const int t = threadIdx.x;
    if (t < s.n && t != s.me) {
        sig_spin_cas(s.peer[t] + blockIdx.x * BG_MAX_PEERS + s.me, 0u, 1u, true, s);
        sig_spin_cas(s.local + blockIdx.x * BG_MAX_PEERS + t, 1u, 0u, false, s);
    }
    if (kSyncAfter) __syncthreads();
}

static int make_sig(bg_ctx* c, int gid, int lane, Sig* s, const Group** gout) {
    if (!c) return fail(BG_EINVAL, "null ctx");
    if (gid < 0 || gid >= (int)c->groups.size()) return fail(BG_EGROUP, "bad gid %d", gid);
    if (lane < 0 || lane >= BG_LANES) return fail(BG_EINVAL, "bad lane %d", lane);
    const Group& g = c->groups[gid];
    s->me = g.me; s->n = g.n;
    s->timeout_ns = (unsigned long long)g_tun.timeout_ms * 1000000ull;
    s->err = c->err_dev;
    s->local = nullptr;
    for (int i = 0; i < BG_MAX_PEERS; ++i) s->peer[i] = nullptr;
    if (g.n > 1) {
        size_t off = (size_t)g.slot * kSlotBytes + (size_t)lane * BG_MAX_CHANNELS * BG_MAX_PEERS * sizeof(uint32_t);
        for (int i = 0; i < g.n; ++i) {
            char* base = c->peer_base[g.ranks[i]];
            if (!base) return fail(BG_ENOTMAPPED, "arena of rank %d is not mapped (bg_arena_import)", g.ranks[i]);
            s->peer[i] = (uint32_t*)(base + off);
        }
        s->local = s->peer[g.me];
    }
    if (gout) *gout = &g;
    return BG_OK;
}

static int resolve(bg_ctx* c, const Group& g, const size_t* offs, size_t bytes, PeerPtrs* out) {
    if (!offs) return fail(BG_EINVAL, "null symmetric-offset array");
    for (int i = 0; i < BG_MAX_PEERS; ++i) out->p[i] = nullptr;
    for (int i = 0; i < g.n; ++i) {
        char* base = c->peer_base[g.ranks[i]];
        if (!base) return fail(BG_ENOTMAPPED, "arena of rank %d is not mapped", g.ranks[i]);
        if (offs[i] % 16) return fail(BG_EINVAL, "symmetric offset %zu not 16-B aligned", offs[i]);
        if (offs[i] < c->pad_bytes || offs[i] + bytes > c->arena_bytes)
            return fail(BG_EINVAL, "symmetric buffer [%zu,+%zu) outside arena", offs[i], bytes);
        out->p[i] = base + offs[i];
    }
    return BG_OK;
}

static int comm_grid(size_t work_items, int threads, int n) {
    long long cap = n == 1 ? g_tun.local_ctas : g_tun.comm_ctas;
    long long want = (long long)((work_items + threads - 1) / threads);
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

__global__ void barrier_kernel(Sig s) { sync_peers<true, true, true>(s); }

extern "C" int bg_barrier(bg_ctx_t c, int gid, int lane, void* stream) {
    Sig s;
    int rc = make_sig(c, gid, lane, &s, nullptr);
    if (rc) return rc;
    BG_CUDA(cudaSetDevice(c->device));
    if (s.n == 1) return BG_OK;
    barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// C1: all-gather (push) fused with cast
// ------------------------------------------------------------------------------------------------
// 256-thread CTAs with <= 128 registers: a communication CTA fits on an SM NEXT TO a persistent GEMM CTA (256 thr x 152 regs,
// 225 KB smem), so collectives on side streams overlap the math instead of queueing behind it.
constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kPullUnroll = 2;   // 16-B vectors per thread per iteration in the pull kernels (x up to 8 peers in flight)

template <typename SrcT, typename DstT>
struct Cvt;
template <> struct Cvt<float, __nv_bfloat16> {  // 8 elements: 32 B in, 16 B out
    static constexpr int kElems = 8;
    __device__ static void load(const float* src, size_t i, uint4* regs) {
        regs[0] = ld16_stream(src + i);
        regs[1] = ld16_stream(src + i + 4);
    }
    __device__ static uint4 convert(const uint4* regs) {
        uint4 o;
        o.x = f2_to_bf2(__uint_as_float(regs[0].x), __uint_as_float(regs[0].y));
        o.y = f2_to_bf2(__uint_as_float(regs[0].z), __uint_as_float(regs[0].w));
        o.z = f2_to_bf2(__uint_as_float(regs[1].x), __uint_as_float(regs[1].y));
        o.w = f2_to_bf2(__uint_as_float(regs[1].z), __uint_as_float(regs[1].w));
        return o;
    }
    static constexpr int kRegs = 2;
};
template <> struct Cvt<__nv_bfloat16, __nv_bfloat16> {
    static constexpr int kElems = 8;
    static constexpr int kRegs = 1;
    __device__ static void load(const __nv_bfloat16* src, size_t i, uint4* regs) { regs[0] = ld16_stream(src + i); }
    __device__ static uint4 convert(const uint4* regs) { return regs[0]; }
};
template <> struct Cvt<float, float> {
    static constexpr int kElems = 4;
    static constexpr int kRegs = 1;
    __device__ static void load(const float* src, size_t i, uint4* regs) { regs[0] = ld16_stream(src + i); }
    __device__ static uint4 convert(const uint4* regs) { return regs[0]; }
};

template <typename SrcT, typename DstT>
__global__ void __launch_bounds__(kThreads, 2) all_gather_push_kernel(PeerPtrs dst, const SrcT* __restrict__ src,
                                                                   size_t shard_elems, Sig s) {
    using C = Cvt<SrcT, DstT>;
    sync_peers<false, false, true>(s);  // every member has finished consuming its dst (it reached this kernel)
    const size_t nvec = shard_elems / C::kElems;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t dst_base = (size_t)s.me * shard_elems * sizeof(DstT);
    for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * kUnroll) {
        uint4 regs[kUnroll][C::kRegs];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            size_t v = v0 + u * stride;
            if (v < nvec) C::load(src, v * C::kElems, regs[u]);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            size_t v = v0 + u * stride;
            if (v < nvec) {
                uint4 o = C::convert(regs[u]);
                for (int k = 0; k < s.n; ++k) {
                    int p = s.me + k; if (p >= s.n) p -= s.n;  // stagger targets across senders
                    st16(dst.p[p] + dst_base + v * 16, o);
                }
            }
        }
    }
    sync_peers<true, true, false>(s);  // my stores are visible everywhere; everyone's shard has landed here
}

extern "C" int bg_all_gather_cast(bg_ctx_t c, int gid, int lane, const void* src, int src_dtype, const size_t* dst_offs,
                                  int dst_dtype, size_t shard_elems, void* stream) {
    Sig s; const Group* g;
    int rc = make_sig(c, gid, lane, &s, &g);
    if (rc) return rc;
    const size_t dsz = dst_dtype == BG_BF16 ? 2 : 4;
    const int per = (src_dtype == BG_F32 && dst_dtype == BG_F32) ? 4 : 8;
    if (shard_elems % per) return fail(BG_EINVAL, "shard_elems %zu must be a multiple of %d (pad the flat buffer)", shard_elems, per);
    if ((uintptr_t)src % 16) return fail(BG_EINVAL, "src not 16-B aligned");
    PeerPtrs dst;
    rc = resolve(c, *g, dst_offs, shard_elems * g->n * dsz, &dst);
    if (rc) return rc;
    if (shard_elems == 0) return BG_OK;
    BG_CUDA(cudaSetDevice(c->device));
    int grid = comm_grid(shard_elems / per / kUnroll + 1, kThreads, g->n);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == BG_F32 && dst_dtype == BG_BF16)
        all_gather_push_kernel<float, __nv_bfloat16><<<grid, kThreads, 0, st>>>(dst, (const float*)src, shard_elems, s);
    else if (src_dtype == BG_BF16 && dst_dtype == BG_BF16)
        all_gather_push_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kThreads, 0, st>>>(dst, (const __nv_bfloat16*)src, shard_elems, s);
    else if (src_dtype == BG_F32 && dst_dtype == BG_F32)
        all_gather_push_kernel<float, float><<<grid, kThreads, 0, st>>>(dst, (const float*)src, shard_elems, s);
    else
        return fail(BG_EUNSUPPORTED, "all_gather_cast %d->%d", src_dtype, dst_dtype);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// C2: reduce-scatter (pull) fused with prescale/postscale, cast and accumulate
// ------------------------------------------------------------------------------------------------
template <bool kSrcBf16>
__device__ __forceinline__ void rs_accumulate(const uint4& v, float* acc, float w) {
    if (kSrcBf16) {
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(f[i], w, acc[i]);
    } else {
        acc[0] = fmaf(__uint_as_float(v.x), w, acc[0]); acc[1] = fmaf(__uint_as_float(v.y), w, acc[1]);
        acc[2] = fmaf(__uint_as_float(v.z), w, acc[2]); acc[3] = fmaf(__uint_as_float(v.w), w, acc[3]);
    }
}

template <bool kSrcBf16, bool kDstBf16>
__global__ void __launch_bounds__(kThreads, 2) reduce_scatter_pull_kernel(PeerPtrs src, void* __restrict__ dst,
                                                                           size_t shard_elems, float prescale,
                                                                           float postscale, int accumulate, Sig s) {
    constexpr int E = kSrcBf16 ? 8 : 4;  // elements per 16-B source vector
    constexpr int U = kPullUnroll;
    sync_peers<false, false, true>(s);  // every member's src is complete (its producer kernels finished before this one)
    const size_t nvec = shard_elems / E;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t slice_off = (size_t)s.me * shard_elems * (kSrcBf16 ? 2 : 4);
    for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * U) {
        uint4 in[U][BG_MAX_PEERS];
        // issue every peer load of this iteration before consuming any: U x (n-1) 16-B NVLink loads in flight per thread
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
#pragma unroll
            for (int p = 0; p < BG_MAX_PEERS; ++p) {
                if (p < s.n && v < nvec) {
                    const char* a = src.p[p] + slice_off + v * 16;
                    in[u][p] = (p == s.me) ? ld16_stream(a) : ld16_peer(a);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
            if (v >= nvec) break;
            float acc[E];
#pragma unroll
            for (int i = 0; i < E; ++i) acc[i] = 0.f;
            // fixed summation order (group order 0..n-1): run-to-run deterministic.  Each rank's contribution is
            // scaled by `prescale` before the sum, as the reference pre-divides (_runtime_utils.py:852).
#pragma unroll
            for (int p = 0; p < BG_MAX_PEERS; ++p)
                if (p < s.n) rs_accumulate<kSrcBf16>(in[u][p], acc, prescale);
#pragma unroll
            for (int i = 0; i < E; ++i) acc[i] *= postscale;
            if (kDstBf16) {
                static_assert(!kDstBf16 || kSrcBf16, "bf16 dst needs bf16 src");
                uint4* d = reinterpret_cast<uint4*>(dst) + v;
                if (accumulate) {
                    float old[8];
                    unpack8(*d, old);
#pragma unroll
                    for (int i = 0; i < E; ++i) acc[i] += old[i];
                }
                st16(d, pack8(acc));
            } else {
                float4* d = reinterpret_cast<float4*>(dst) + v * (E / 4);
#pragma unroll
                for (int q = 0; q < E / 4; ++q) {
                    float4 o = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
                    if (accumulate) {
                        float4 old = d[q];
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    d[q] = o;
                }
            }
        }
    }
    sync_peers<true, false, false>(s);  // every member has finished reading my src: it may be overwritten
}

extern "C" int bg_reduce_scatter_acc(bg_ctx_t c, int gid, int lane, const size_t* src_offs, int src_dtype, void* dst,
                                     int dst_dtype, size_t shard_elems, float prescale, float postscale, int accumulate,
                                     void* stream) {
    Sig s; const Group* g;
    int rc = make_sig(c, gid, lane, &s, &g);
    if (rc) return rc;
    const int per = src_dtype == BG_BF16 ? 8 : 4;
    if (shard_elems % per) return fail(BG_EINVAL, "shard_elems %zu must be a multiple of %d", shard_elems, per);
    if ((uintptr_t)dst % 16) return fail(BG_EINVAL, "dst not 16-B aligned");
    PeerPtrs src;
    rc = resolve(c, *g, src_offs, shard_elems * g->n * (src_dtype == BG_BF16 ? 2 : 4), &src);
    if (rc) return rc;
    if (shard_elems == 0) return BG_OK;
    BG_CUDA(cudaSetDevice(c->device));
    int grid = comm_grid(shard_elems / per / kPullUnroll + 1, kThreads, g->n);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == BG_BF16 && dst_dtype == BG_F32)
        reduce_scatter_pull_kernel<true, false><<<grid, kThreads, 0, st>>>(src, dst, shard_elems, prescale, postscale, accumulate, s);
    else if (src_dtype == BG_BF16 && dst_dtype == BG_BF16)
        reduce_scatter_pull_kernel<true, true><<<grid, kThreads, 0, st>>>(src, dst, shard_elems, prescale, postscale, accumulate, s);
    else if (src_dtype == BG_F32 && dst_dtype == BG_F32)
        reduce_scatter_pull_kernel<false, false><<<grid, kThreads, 0, st>>>(src, dst, shard_elems, prescale, postscale, accumulate, s);
    else
        return fail(BG_EUNSUPPORTED, "reduce_scatter %d->%d", src_dtype, dst_dtype);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// C2 + optimizer: reduce-scatter whose epilogue IS the AdamW step on the fp32 shard.  The reduced gradient
// never touches HBM: g = sum_p(G_p[slice]) * prescale * postscale stays in registers and updates (param, exp_avg, exp_avg_sq).
// Same update rule as torch.optim.AdamW / apex FusedAdam(adam_w_mode=True) (galvatron/core/runtime/utils.py:137-150).
// ------------------------------------------------------------------------------------------------
struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2_sqrt;
};

template <bool kSrcBf16>
__global__ void __launch_bounds__(kThreads, 2) reduce_scatter_adamw_kernel(PeerPtrs src, float* __restrict__ param,
                                                                         float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                                         size_t shard_elems, float prescale, float postscale,
                                                                         AdamArgs a, Sig s) {
    constexpr int E = kSrcBf16 ? 8 : 4;
    sync_peers<false, false, true>(s);
    const size_t nvec = shard_elems / E;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t slice_off = (size_t)s.me * shard_elems * (kSrcBf16 ? 2 : 4);
    const float step_size = a.lr / a.bias_corr1, decay = 1.f - a.lr * a.weight_decay;
    constexpr int U = kPullUnroll;
    for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * U) {
        uint4 in[U][BG_MAX_PEERS];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
#pragma unroll
            for (int p = 0; p < BG_MAX_PEERS; ++p)
                if (p < s.n && v < nvec) {
                    const char* ad = src.p[p] + slice_off + v * 16;
                    in[u][p] = (p == s.me) ? ld16_stream(ad) : ld16_peer(ad);
                }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
            if (v >= nvec) break;
            float g[E];
#pragma unroll
            for (int i = 0; i < E; ++i) g[i] = 0.f;
#pragma unroll
            for (int p = 0; p < BG_MAX_PEERS; ++p)
                if (p < s.n) rs_accumulate<kSrcBf16>(in[u][p], g, prescale);
            float4* pp = reinterpret_cast<float4*>(param) + v * (E / 4);
            float4* pm = reinterpret_cast<float4*>(exp_avg) + v * (E / 4);
            float4* pv = reinterpret_cast<float4*>(exp_avg_sq) + v * (E / 4);
#pragma unroll
            for (int q = 0; q < E / 4; ++q) {
                float4 w = pp[q], m = pm[q], vv = pv[q];
                float* wf = reinterpret_cast<float*>(&w); float* mf = reinterpret_cast<float*>(&m); float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float gi = g[4 * q + i] * postscale;
                    mf[i] = a.beta1 * mf[i] + (1.f - a.beta1) * gi;
                    vf[i] = a.beta2 * vf[i] + (1.f - a.beta2) * gi * gi;
                    const float denom = sqrtf(vf[i]) / a.bias_corr2_sqrt + a.eps;
                    wf[i] = wf[i] * decay - step_size * mf[i] / denom;
                }
                pp[q] = w; pm[q] = m; pv[q] = vv;
            }
        }
    }
    sync_peers<true, false, false>(s);
}

extern "C" int bg_reduce_scatter_adamw(bg_ctx_t c, int gid, int lane, const size_t* src_offs, int src_dtype, float* param,
                                       float* exp_avg, float* exp_avg_sq, size_t shard_elems, float prescale, float postscale,
                                       float lr, float beta1, float beta2, float eps, float weight_decay, long long step,
                                       void* stream) {
    Sig s; const Group* g;
    int rc = make_sig(c, gid, lane, &s, &g);
    if (rc) return rc;
    const int per = src_dtype == BG_BF16 ? 8 : 4;
    if (shard_elems % per) return fail(BG_EINVAL, "shard_elems %zu must be a multiple of %d", shard_elems, per);
    if (((uintptr_t)param | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) return fail(BG_EINVAL, "optimizer state not 16-B aligned");
    if (step < 1) return fail(BG_EINVAL, "adam step must be >= 1");
    PeerPtrs src;
    rc = resolve(c, *g, src_offs, shard_elems * g->n * (src_dtype == BG_BF16 ? 2 : 4), &src);
    if (rc) return rc;
    if (shard_elems == 0) return BG_OK;
    BG_CUDA(cudaSetDevice(c->device));
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.bias_corr1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias_corr2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    int grid = comm_grid(shard_elems / per / kPullUnroll + 1, kThreads, g->n);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == BG_BF16) reduce_scatter_adamw_kernel<true><<<grid, kThreads, 0, st>>>(src, param, exp_avg, exp_avg_sq, shard_elems, prescale, postscale, a, s);
    else reduce_scatter_adamw_kernel<false><<<grid, kThreads, 0, st>>>(src, param, exp_avg, exp_avg_sq, shard_elems, prescale, postscale, a, s);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// C3/C5/C6/C13: all-reduce, one-shot (small) and two-shot (large)
// ------------------------------------------------------------------------------------------------
template <bool kBf16, bool kMax>
__device__ __forceinline__ void ar_combine(const uint4& v, float* acc, bool first) {
    constexpr int E = kBf16 ? 8 : 4;
    float f[E];
    if (kBf16) unpack8(v, f);
    else { f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w); }
#pragma unroll
    for (int i = 0; i < E; ++i) acc[i] = first ? f[i] : (kMax ? fmaxf(acc[i], f[i]) : acc[i] + f[i]);
}

template <bool kBf16>
__device__ __forceinline__ uint4 ar_pack(const float* acc, float scale) {
    if (kBf16) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = acc[i] * scale;
        return pack8(t);
    }
    uint4 o;
    o.x = __float_as_uint(acc[0] * scale); o.y = __float_as_uint(acc[1] * scale);
    o.z = __float_as_uint(acc[2] * scale); o.w = __float_as_uint(acc[3] * scale);
    return o;
}

// one-shot: every member reads all n buffers in full
template <bool kBf16, bool kMax>
__global__ void __launch_bounds__(kThreads, 2) all_reduce_oneshot_kernel(PeerPtrs src, void* __restrict__ dst, size_t nvec,
                                                                       float scale, Sig s) {
    constexpr int E = kBf16 ? 8 : 4;
    sync_peers<false, false, true>(s);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        uint4 in[BG_MAX_PEERS];
#pragma unroll
        for (int p = 0; p < BG_MAX_PEERS; ++p)
            if (p < s.n) in[p] = (p == s.me) ? ld16_stream(src.p[p] + v * 16) : ld16_peer(src.p[p] + v * 16);
        float acc[E];
#pragma unroll
        for (int p = 0; p < BG_MAX_PEERS; ++p)
            if (p < s.n) ar_combine<kBf16, kMax>(in[p], acc, p == 0);
        st16(reinterpret_cast<uint4*>(dst) + v, ar_pack<kBf16>(acc, scale));
    }
    sync_peers<true, false, false>(s);
}

// two-shot: reduce my slice into my own src (peer-visible), barrier, gather every member's reduced slice.
// Vector v of a slice is always handled by the same (CTA, thread) on every member, so the per-CTA channel
// barrier between the two phases is sufficient.
template <bool kBf16, bool kMax>
__global__ void __launch_bounds__(kThreads, 2) all_reduce_twoshot_kernel(PeerPtrs src, void* __restrict__ dst,
                                                                          size_t slice_vec, float scale, Sig s) {
    constexpr int E = kBf16 ? 8 : 4;
    constexpr int U = kPullUnroll;
    sync_peers<false, false, true>(s);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t my0 = (size_t)s.me * slice_vec;
    for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < slice_vec; v0 += stride * U) {
        uint4 in[U][BG_MAX_PEERS];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
#pragma unroll
            for (int p = 0; p < BG_MAX_PEERS; ++p)
                if (p < s.n && v < slice_vec)
                    in[u][p] = (p == s.me) ? ld16_stream(src.p[p] + (my0 + v) * 16) : ld16_peer(src.p[p] + (my0 + v) * 16);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
            if (v >= slice_vec) break;
            float acc[E];
#pragma unroll
            for (int p = 0; p < BG_MAX_PEERS; ++p)
                if (p < s.n) ar_combine<kBf16, kMax>(in[u][p], acc, p == 0);
            uint4 o = ar_pack<kBf16>(acc, scale);
            st16(src.p[s.me] + (my0 + v) * 16, o);
            st16(reinterpret_cast<uint4*>(dst) + my0 + v, o);
        }
    }
    sync_peers<true, true, true>(s);
    // gather every member's reduced slice (vector v of a slice is handled by the same CTA on every member)
    for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < slice_vec; v0 += stride * U) {
        uint4 in[U][BG_MAX_PEERS];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
#pragma unroll
            for (int k = 1; k < BG_MAX_PEERS; ++k)
                if (k < s.n && v < slice_vec) {
                    int p = s.me + k; if (p >= s.n) p -= s.n;
                    in[u][k] = ld16_peer(src.p[p] + ((size_t)p * slice_vec + v) * 16);
                }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * stride;
            if (v >= slice_vec) break;
#pragma unroll
            for (int k = 1; k < BG_MAX_PEERS; ++k)
                if (k < s.n) {
                    int p = s.me + k; if (p >= s.n) p -= s.n;
                    st16(reinterpret_cast<uint4*>(dst) + (size_t)p * slice_vec + v, in[u][k]);
                }
        }
    }
    sync_peers<true, false, false>(s);
}

extern "C" int bg_all_reduce(bg_ctx_t c, int gid, int lane, const size_t* src_offs, void* dst, size_t elems, int dtype,
                             int redop, float scale, void* stream) {
    Sig s; const Group* g;
    int rc = make_sig(c, gid, lane, &s, &g);
    if (rc) return rc;
    if (dtype != BG_BF16 && dtype != BG_F32) return fail(BG_EUNSUPPORTED, "all_reduce dtype %d", dtype);
    if (redop != BG_SUM && redop != BG_MAX) return fail(BG_EUNSUPPORTED, "all_reduce op %d", redop);
    const int per = dtype == BG_BF16 ? 8 : 4;
    const size_t esz = dtype == BG_BF16 ? 2 : 4;
    if (elems % per) return fail(BG_EINVAL, "all_reduce elems %zu must be a multiple of %d (pad)", elems, per);
    if ((uintptr_t)dst % 16) return fail(BG_EINVAL, "dst not 16-B aligned");
    PeerPtrs src;
    rc = resolve(c, *g, src_offs, elems * esz, &src);
    if (rc) return rc;
    if (elems == 0) return BG_OK;
    BG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nvec = elems / per;
    const bool twoshot = g->n > 1 && elems * esz > (size_t)g_tun.oneshot_bytes && nvec % g->n == 0;
    const bool bf = dtype == BG_BF16, mx = redop == BG_MAX;
#define BG_AR_DISPATCH(KERNEL, NV)                                                                    \
    do {                                                                                              \
        int grid = comm_grid((NV) / kPullUnroll + 1, kThreads, g->n);                                                   \
        if (bf && !mx) KERNEL<true, false><<<grid, kThreads, 0, st>>>(src, dst, (NV), scale, s);      \
        else if (bf && mx) KERNEL<true, true><<<grid, kThreads, 0, st>>>(src, dst, (NV), scale, s);   \
        else if (!bf && !mx) KERNEL<false, false><<<grid, kThreads, 0, st>>>(src, dst, (NV), scale, s); \
        else KERNEL<false, true><<<grid, kThreads, 0, st>>>(src, dst, (NV), scale, s);                \
    } while (0)
    if (twoshot) BG_AR_DISPATCH(all_reduce_twoshot_kernel, nvec / g->n);
    else BG_AR_DISPATCH(all_reduce_oneshot_kernel, nvec);
#undef BG_AR_DISPATCH
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// C10: Ulysses all-to-all fused with the head/seq transpose (pull; up to 4 tensors per launch)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxA2A = 4;
struct A2ADev {
    PeerPtrs src;
    char* dst;
    long long batch, rows, row_vec;           // row_vec = 16-B vectors per row
    long long src_bs, src_rs, src_me_off;     // in 16-B vectors
    long long dst_bs, dst_rs, dst_peer_off;   // in 16-B vectors
    long long total_vec;                      // batch * rows * row_vec * n
};
struct A2AArgs {
    A2ADev t[kMaxA2A];
    int n_tensors;
};

__global__ void __launch_bounds__(kThreads, 2) all_to_all_rows_kernel(A2AArgs a, Sig s) {
    sync_peers<false, false, true>(s);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int ti = 0; ti < a.n_tensors; ++ti) {
        const A2ADev& d = a.t[ti];
        const long long per_peer = d.batch * d.rows * d.row_vec;
        for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < (size_t)d.total_vec; i0 += stride * kUnroll) {
            uint4 regs[kUnroll];
            long long dsts[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                size_t i = i0 + u * stride;
                dsts[u] = -1;
                if (i < (size_t)d.total_vec) {
                    int k = (int)(i / per_peer);
                    long long r = (long long)(i - (size_t)k * per_peer);
                    int q = s.me + k; if (q >= s.n) q -= s.n;
                    long long c = r % d.row_vec; r /= d.row_vec;
                    long long row = r % d.rows, b = r / d.rows;
                    const char* sp = d.src.p[q] + (b * d.src_bs + row * d.src_rs + (long long)s.me * d.src_me_off + c) * 16;
                    regs[u] = (q == s.me) ? ld16_stream(sp) : ld16_peer(sp);
                    dsts[u] = b * d.dst_bs + row * d.dst_rs + (long long)q * d.dst_peer_off + c;
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (dsts[u] >= 0) st16(d.dst + dsts[u] * 16, regs[u]);
        }
    }
    sync_peers<true, false, false>(s);
}

extern "C" int bg_all_to_all_rows(bg_ctx_t c, int gid, int lane, const bg_a2a_desc* descs, int n_descs, int dtype,
                                  void* stream) {
    Sig s; const Group* g;
    int rc = make_sig(c, gid, lane, &s, &g);
    if (rc) return rc;
    if (!descs || n_descs < 1 || n_descs > kMaxA2A) return fail(BG_EINVAL, "1..%d tensors per all_to_all launch", kMaxA2A);
    const long long esz = dtype == BG_BF16 ? 2 : 4, per = 16 / esz;
    A2AArgs a;
    a.n_tensors = n_descs;
    size_t max_vec = 0;
    for (int i = 0; i < n_descs; ++i) {
        const bg_a2a_desc& d = descs[i];
        if (d.row_elems % per || d.src_bs % per || d.src_rs % per || d.src_me_off % per || d.dst_bs % per ||
            d.dst_rs % per || d.dst_peer_off % per)
            return fail(BG_EINVAL, "all_to_all: strides/row length must be multiples of %lld elements", per);
        if ((uintptr_t)d.dst % 16) return fail(BG_EINVAL, "all_to_all dst not 16-B aligned");
        // extent of the peer's source that may be touched
        long long span = (d.batch - 1) * d.src_bs + (d.rows - 1) * d.src_rs + (long long)(g->n - 1) * d.src_me_off + d.row_elems;
        rc = resolve(c, *g, d.src_offs, (size_t)span * esz, &a.t[i].src);
        if (rc) return rc;
        a.t[i].dst = (char*)d.dst;
        a.t[i].batch = d.batch; a.t[i].rows = d.rows; a.t[i].row_vec = d.row_elems / per;
        a.t[i].src_bs = d.src_bs / per; a.t[i].src_rs = d.src_rs / per; a.t[i].src_me_off = d.src_me_off / per;
        a.t[i].dst_bs = d.dst_bs / per; a.t[i].dst_rs = d.dst_rs / per; a.t[i].dst_peer_off = d.dst_peer_off / per;
        a.t[i].total_vec = d.batch * d.rows * a.t[i].row_vec * g->n;
        if ((size_t)a.t[i].total_vec > max_vec) max_vec = (size_t)a.t[i].total_vec;
    }
    BG_CUDA(cudaSetDevice(c->device));
    int grid = comm_grid(max_vec / kUnroll + 1, kThreads, g->n);
    all_to_all_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(a, s);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// C11: pipeline p2p -- peer copy on the caller's (side) stream + device flags, no device-wide sync.
// Flags in every arena: F[other_rank][flag_id][2]; [0] "a message from other has landed here",
// [1] "other has consumed the message I sent".  A sender re-uses a slot only after the receiver's ack.
// ------------------------------------------------------------------------------------------------
__global__ void p2p_raise_kernel(uint32_t* flag, unsigned long long timeout_ns, int* err) {
    __threadfence_system();
    Sig s; s.timeout_ns = timeout_ns; s.err = err;
    sig_spin_cas(flag, 0u, 1u, true, s);
}
__global__ void p2p_consume_kernel(uint32_t* flag, unsigned long long timeout_ns, int* err) {
    Sig s; s.timeout_ns = timeout_ns; s.err = err;
    sig_spin_cas(flag, 1u, 0u, false, s);
}

static uint32_t* p2p_flag(bg_ctx* c, int owner_rank, int other_rank, int flag_id, int which) {
    char* base = c->peer_base[owner_rank];
    if (!base) return nullptr;
    size_t off = c->slot_of.size() * kSlotBytes + (((size_t)other_rank * kP2PFlags + flag_id) * 2 + which) * sizeof(uint32_t);
    return (uint32_t*)(base + off);
}

static int p2p_args(bg_ctx* c, int peer, int flag_id, const char* who) {
    if (!c || peer < 0 || peer >= c->world || peer == c->rank || flag_id < 0 || flag_id >= kP2PFlags)
        return fail(BG_EINVAL, "%s: bad peer/flag %d/%d", who, peer, flag_id);
    if (!c->peer_base[peer]) return fail(BG_ENOTMAPPED, "arena of rank %d is not mapped", peer);
    return BG_OK;
}

extern "C" int bg_p2p_send(bg_ctx_t c, int peer, size_t dst_off, const void* src, size_t bytes, int flag_id, void* stream) {
    int rc = p2p_args(c, peer, flag_id, "bg_p2p_send");
    if (rc) return rc;
    if (dst_off < c->pad_bytes || dst_off + bytes > c->arena_bytes) return fail(BG_EINVAL, "p2p destination outside arena");
    BG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long to = (unsigned long long)g_tun.timeout_ms * 1000000ull;
    bool first;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        first = c->p2p_sent[peer][flag_id]++ == 0;
    }
    if (!first) {  // the receiver must have released the slot (bg_p2p_release) before it is overwritten
        p2p_consume_kernel<<<1, 1, 0, st>>>(p2p_flag(c, c->rank, peer, flag_id, 1), to, c->err_dev);
        BG_CHECK_LAUNCH();
    }
    if (bytes) BG_CUDA(cudaMemcpyAsync(c->peer_base[peer] + dst_off, src, bytes, cudaMemcpyDeviceToDevice, st));
    p2p_raise_kernel<<<1, 1, 0, st>>>(p2p_flag(c, peer, c->rank, flag_id, 0), to, c->err_dev);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_p2p_wait(bg_ctx_t c, int peer, int flag_id, void* stream) {
    int rc = p2p_args(c, peer, flag_id, "bg_p2p_wait");
    if (rc) return rc;
    BG_CUDA(cudaSetDevice(c->device));
    p2p_consume_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p2p_flag(c, c->rank, peer, flag_id, 0),
                                                          (unsigned long long)g_tun.timeout_ms * 1000000ull, c->err_dev);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_p2p_release(bg_ctx_t c, int peer, int flag_id, void* stream) {
    int rc = p2p_args(c, peer, flag_id, "bg_p2p_release");
    if (rc) return rc;
    BG_CUDA(cudaSetDevice(c->device));
    p2p_raise_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p2p_flag(c, peer, c->rank, flag_id, 1),
                                                        (unsigned long long)g_tun.timeout_ms * 1000000ull, c->err_dev);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// C5/C8 fused: GEMM whose epilogue reduce-scatters over the group (tcgen05 tiles -> peer HBM -> tile reducer)
// ------------------------------------------------------------------------------------------------
int bg_gemm_scatter_launch(const void* a, const void* b, long long m, long long n, long long k, int layout, int p, int me,
                           void* const* partial_ptrs, uint32_t* const* flag_ptrs, void* out, unsigned long long timeout_ns,
                           int* err_dev, cudaStream_t st);

extern "C" int bg_gemm_reduce_scatter(bg_ctx_t c, int gid, int lane, const void* a, const void* b, long long m, long long n,
                                      long long k, int layout, const size_t* partial_offs, const size_t* flag_offs, void* out,
                                      void* stream) {
    Sig s; const Group* g;
    int rc = make_sig(c, gid, lane, &s, &g);
    if (rc) return rc;
    if (g->n < 2) return fail(BG_EINVAL, "bg_gemm_reduce_scatter needs a group of >= 2 ranks (use bg_gemm_bf16)");
    PeerPtrs partial, flags;
    rc = resolve(c, *g, partial_offs, (size_t)m * n * 2, &partial);
    if (rc) return rc;
    const size_t n_flags = (size_t)((m / g->n + 127) / 128) * ((n + 255) / 256);
    rc = resolve(c, *g, flag_offs, n_flags * sizeof(uint32_t), &flags);
    if (rc) return rc;
    BG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    // Entry barrier: every member's previous use of the partial buffers and counters (its last reducer, earlier in this same
    // stream) has drained before any peer may store into them again.
    barrier_kernel<<<1, 32, 0, st>>>(s);
    BG_CHECK_LAUNCH();
    void* pp[BG_MAX_PEERS]; uint32_t* fp[BG_MAX_PEERS];
    for (int i = 0; i < BG_MAX_PEERS; ++i) { pp[i] = partial.p[i]; fp[i] = (uint32_t*)flags.p[i]; }
    return bg_gemm_scatter_launch(a, b, m, n, k, layout, g->n, g->me, pp, fp, out, (unsigned long long)g_tun.timeout_ms * 1000000ull,
                                  c->err_dev, st);
}

// ---- NVLS: multicast object over one group's symmetric buffer ---------------------------------------------------------
static int mc_group(bg_ctx* c, int gid, const Group** gout) {
    if (!c) return fail(BG_EINVAL, "null ctx");
    if (gid < 0 || gid >= (int)c->groups.size()) return fail(BG_EGROUP, "bad gid %d", gid);
    if (!c->vmm || !c->mc_supported) return fail(BG_EINVAL, "NVLS needs a BG_CTX_VMM context on a multicast-capable device");
    *gout = &c->groups[gid];
    if ((*gout)->n < 2) return fail(BG_EINVAL, "NVLS needs a group of >= 2 ranks");
    return BG_OK;
}

extern "C" int bg_group_mc_create(bg_ctx_t c, int gid, size_t bytes, int* fd_out) {
    const Group* g;
    int rc = mc_group(c, gid, &g);
    if (rc) return rc;
    if (!fd_out || !bytes) return fail(BG_EINVAL, "null arg");
    BG_CUDA(cudaSetDevice(c->device));
    bg_ctx::McGroup& m = c->mc_of[gid];
    if (m.mc) return fail(BG_EINVAL, "group %d already has a multicast object", gid);
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned)g->n;
    mp.size = (bytes + c->mc_gran - 1) / c->mc_gran * c->mc_gran;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    BG_DRV(drv().MulticastCreate(&m.mc, &mp));
    m.bytes = mp.size;
    BG_DRV(drv().MemExportToShareableHandle(fd_out, m.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    return BG_OK;
}

// fd >= 0: import the creator's object (every other member); then add this rank's device.  ALL members must have joined
// before any of them binds (host-side barrier).
extern "C" int bg_group_mc_join(bg_ctx_t c, int gid, int fd, size_t bytes) {
    const Group* g;
    int rc = mc_group(c, gid, &g);
    if (rc) return rc;
    BG_CUDA(cudaSetDevice(c->device));
    bg_ctx::McGroup& m = c->mc_of[gid];
    if (fd >= 0) {
        if (m.mc) return fail(BG_EINVAL, "group %d already has a multicast object", gid);
        BG_DRV(drv().MemImportFromShareableHandle(&m.mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
        m.bytes = (bytes + c->mc_gran - 1) / c->mc_gran * c->mc_gran;
    }
    if (!m.mc) return fail(BG_EINVAL, "group %d has no multicast object (create or import first)", gid);
    CUdevice dev = 0;
    BG_DRV(drv().DeviceGet(&dev, c->device));
    BG_DRV(drv().MulticastAddDevice(m.mc, dev));
    return BG_OK;
}

extern "C" int bg_group_mc_bind(bg_ctx_t c, int gid, size_t arena_offset) {
    const Group* g;
    int rc = mc_group(c, gid, &g);
    if (rc) return rc;
    BG_CUDA(cudaSetDevice(c->device));
    bg_ctx::McGroup& m = c->mc_of[gid];
    if (!m.mc || m.bound) return fail(BG_EINVAL, "group %d: multicast object missing or already bound", gid);
    if (arena_offset % c->mc_gran || arena_offset < c->pad_bytes || arena_offset + m.bytes > c->arena_bytes)
        return fail(BG_EINVAL, "NVLS buffer [%zu,+%zu) must be multicast-granularity (%zu) aligned inside the arena", arena_offset, m.bytes,
                    c->mc_gran);
    BG_DRV(drv().MulticastBindMem(m.mc, 0, c->arena_handle, arena_offset, m.bytes, 0));
    rc = map_rw(&m.va, m.bytes, c->mc_gran, m.mc, c->device);
    if (rc) return rc;
    m.arena_off = arena_offset;
    m.bound = true;
    return BG_OK;
}

namespace {

__device__ __forceinline__ uint4 mm_ld_reduce_bf16(const void* mc) {   // sum over every member's copy, fp32 accumulation in the switch
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ float4 mm_ld_reduce_f32(const void* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_16(void* mc, const uint4& v) {    // one store, lands in every member's copy
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}

// Two-shot all-reduce through the switch, in place on the group's multicast-bound buffer, then a local copy to dst.
//   phase 1: member r owns vectors [r*per, (r+1)*per): ld_reduce pulls the SUM of all members' values (one NVLink read of the
//            reduced data instead of p-1 reads), scale, multimem.st pushes the result into every member's buffer
//   phase 2: after the barrier every member's buffer holds the full result; copy it out (local HBM)
// NVLink bytes per GPU: N/p received + N/p sent through the switch's reduction / replication, vs 2(p-1)/p*N for the P2P two-shot.
template <bool kBf16>
__global__ void __launch_bounds__(256, 2) all_reduce_nvls_kernel(char* mc, const char* local, char* dst, size_t vecs, float scale,
                                                                Sig s) {
    sync_peers<false, false, true>(s);   // every member's input is complete (its producers precede this kernel in its stream)
    const size_t per = (vecs + s.n - 1) / s.n;
    const size_t lo = per * s.me, hi = lo + per < vecs ? lo + per : vecs;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int kU = 4;                // in-switch reductions in flight per thread
    for (size_t v0 = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < hi; v0 += stride * kU) {
        uint4 val[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t v = v0 + (size_t)u * stride;
            if (v < hi) {
                if (kBf16) {
                    val[u] = mm_ld_reduce_bf16(mc + v * 16);
                } else {
                    float4 in = mm_ld_reduce_f32(mc + v * 16);
                    val[u] = make_uint4(__float_as_uint(in.x), __float_as_uint(in.y), __float_as_uint(in.z), __float_as_uint(in.w));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const size_t v = v0 + (size_t)u * stride;
            if (v < hi) {
                uint4 out = val[u];
                if (scale != 1.0f) {
                    if (kBf16) {
                        float f[8];
                        unpack8(out, f);
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] *= scale;
                        out = pack8(f);
                    } else {
                        out.x = __float_as_uint(__uint_as_float(out.x) * scale); out.y = __float_as_uint(__uint_as_float(out.y) * scale);
                        out.z = __float_as_uint(__uint_as_float(out.z) * scale); out.w = __float_as_uint(__uint_as_float(out.w) * scale);
                    }
                }
                mm_st_16(mc + v * 16, out);
            }
        }
    }
    sync_peers<true, true, true>(s);     // my stores are visible everywhere and everyone's slice has landed here
    if (dst != nullptr) {
        for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < vecs; v0 += stride * kU) {
            uint4 val[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (v0 + (size_t)u * stride < vecs) val[u] = ld16_stream(local + (v0 + (size_t)u * stride) * 16);
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (v0 + (size_t)u * stride < vecs) st16(dst + (v0 + (size_t)u * stride) * 16, val[u]);
        }
        // the buffer may be refilled by the next call's producers only after every member has finished reading it: the next
        // call's entry barrier cannot give that (it waits for producers, not consumers), so leave through a barrier
        sync_peers<true, false, false>(s);
    }
}

}  // namespace

extern "C" int bg_all_reduce_nvls(bg_ctx_t c, int gid, int lane, size_t byte_offset, void* dst, size_t elems, int dtype, float scale,
                                  void* stream) {
    Sig s; const Group* g;
    int rc = make_sig(c, gid, lane, &s, &g);
    if (rc) return rc;
    auto it = c->mc_of.find(gid);
    if (it == c->mc_of.end() || !it->second.bound) return fail(BG_EINVAL, "group %d has no bound NVLS buffer", gid);
    const bg_ctx::McGroup& m = it->second;
    const size_t esz = dtype == BG_BF16 ? 2 : dtype == BG_F32 ? 4 : 0;
    if (!esz) return fail(BG_EUNSUPPORTED, "bg_all_reduce_nvls: bf16 or fp32");
    if (elems * esz % 16) return fail(BG_EINVAL, "bg_all_reduce_nvls: payload must be a multiple of 16 bytes");
    if (byte_offset % 16 || byte_offset + elems * esz > m.bytes)
        return fail(BG_EINVAL, "bg_all_reduce_nvls: [%zu,+%zu) outside the bound buffer (%zu B) or misaligned", byte_offset, elems * esz, m.bytes);
    BG_CUDA(cudaSetDevice(c->device));
    const size_t vecs = elems * esz / 16;
    const int grid = comm_grid((vecs + g->n - 1) / g->n, 256, g->n);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == BG_BF16)
        all_reduce_nvls_kernel<true><<<grid, 256, 0, st>>>((char*)m.va + byte_offset, c->arena + m.arena_off + byte_offset, (char*)dst, vecs, scale, s);
    else
        all_reduce_nvls_kernel<false><<<grid, 256, 0, st>>>((char*)m.va + byte_offset, c->arena + m.arena_off + byte_offset, (char*)dst, vecs, scale, s);
    BG_CHECK_LAUNCH();
    return BG_OK;
}
