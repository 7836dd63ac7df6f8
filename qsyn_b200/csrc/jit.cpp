// jit.cpp -- run-time compilation and cache of specialised sweep kernels.
//
// request(): hash the sweep program, look it up, on a miss generate CUDA C++ (jit_codegen.cpp) and queue it for a
// pool of compiler threads (NVRTC -> sm_100a cubin; ~0.1 s per sweep, all host cores).  The caller never waits:
// a sweep whose kernel is not ready yet runs in the interpreter (kernels.cu), the next run of the same program
// (operand B of the job, the next column block, the next check of the same circuit) finds it compiled.
// Compiled code is loaded with cuLibraryLoadData (context independent: one load serves every GPU of the process)
// and launched with cuLaunchKernel; both come from the driver through cudaGetDriverEntryPoint, NVRTC through dlopen,
// so libqsx.so links against neither and a host without them simply keeps interpreting.
#include "jit.hpp"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace qsx::jit {

namespace {

// ---- NVRTC (dlopen) ------------------------------------------------------------------------
struct Nvrtc {
    void* handle = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*)                                               = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*)                                                                 = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*)                                                                       = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*)                                                            = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*)                                                                  = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*)                                                                       = nullptr;
    const char* (*GetErrorString)(nvrtcResult)                                                                         = nullptr;
    nvrtcResult (*Version)(int*, int*)                                                                                 = nullptr;
};

// ---- driver entry points (cudaGetDriverEntryPoint) -------------------------------------------
struct Driver {
    CUresult (*LibraryLoadData)(CUlibrary*, const void*, CUjit_option*, void**, unsigned, CUlibraryOption*, void**, unsigned) = nullptr;
    CUresult (*LibraryGetKernel)(CUkernel*, CUlibrary, const char*)                                                           = nullptr;
    CUresult (*KernelSetAttribute)(CUfunction_attribute, int, CUkernel, CUdevice)                                             = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**)                                              = nullptr;  // optional
    CUresult (*GetErrorString)(CUresult, const char**)                                                                        = nullptr;
    CUresult (*CtxGetDevice)(CUdevice*)                                                                                       = nullptr;
};

enum State : int { kQueued = 0, kCompiling = 1, kReady = 2, kFailed = 3 };

}  // namespace

struct Kernel {
    uint64_t h0 = 0, h1 = 0;
    int R = 0, threads = 0, min_ctas = 1;
    bool persistent = false;
    std::string name, source;
    std::vector<unsigned char> program;  // copy of the serialized sweep until the worker has generated the source from it
    std::atomic<int> state{kQueued};
    std::vector<char> cubin;
    // loading (first launch)
    std::mutex load_mu;
    CUlibrary lib  = nullptr;
    CUkernel kern  = nullptr;
    bool load_failed = false;
    bool attr_set[64] = {};
};

namespace {

// The synchronisation objects and containers the (detached, never-ending) compiler threads touch are allocated once
// and never destroyed: a static destructor running at process exit while a worker waits on the condition variable
// would block forever in pthread_cond_destroy.
std::mutex& g_mu                   = *new std::mutex;  // cache, queue, stats
std::condition_variable& g_cv_work = *new std::condition_variable;
std::condition_variable& g_cv_done = *new std::condition_variable;
auto& g_cache                      = *new std::map<std::pair<uint64_t, uint64_t>, std::unique_ptr<Kernel>>;
std::deque<Kernel*>& g_queue       = *new std::deque<Kernel*>;
bool g_pool_started = false;
bool g_stop         = false;  // set at process exit: no new compilations, wait for the running ones
int g_active        = 0;      // workers inside the compiler right now
int g_hook_registrations = 0;
void stop_pool_at_exit();
Stats g_stats;
std::string& g_last_error = *new std::string;

Nvrtc g_nvrtc;
Driver g_drv;
int g_avail = -1;  // -1 unknown, 0 no, 1 yes
std::string& g_why = *new std::string;

bool load_nvrtc(std::string& why) {
    char const* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (char const* nm : names) {
        g_nvrtc.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (g_nvrtc.handle != nullptr) break;
    }
    if (g_nvrtc.handle == nullptr) {
        why = std::string("libnvrtc is not loadable: ") + dlerror();
        return false;
    }
    bool ok  = true;
    auto sym = [&](char const* n) {
        void* p = dlsym(g_nvrtc.handle, n);
        if (p == nullptr) ok = false;
        return p;
    };
    g_nvrtc.CreateProgram     = reinterpret_cast<decltype(g_nvrtc.CreateProgram)>(sym("nvrtcCreateProgram"));
    g_nvrtc.CompileProgram    = reinterpret_cast<decltype(g_nvrtc.CompileProgram)>(sym("nvrtcCompileProgram"));
    g_nvrtc.GetCUBINSize      = reinterpret_cast<decltype(g_nvrtc.GetCUBINSize)>(sym("nvrtcGetCUBINSize"));
    g_nvrtc.GetCUBIN          = reinterpret_cast<decltype(g_nvrtc.GetCUBIN)>(sym("nvrtcGetCUBIN"));
    g_nvrtc.GetProgramLogSize = reinterpret_cast<decltype(g_nvrtc.GetProgramLogSize)>(sym("nvrtcGetProgramLogSize"));
    g_nvrtc.GetProgramLog     = reinterpret_cast<decltype(g_nvrtc.GetProgramLog)>(sym("nvrtcGetProgramLog"));
    g_nvrtc.DestroyProgram    = reinterpret_cast<decltype(g_nvrtc.DestroyProgram)>(sym("nvrtcDestroyProgram"));
    g_nvrtc.GetErrorString    = reinterpret_cast<decltype(g_nvrtc.GetErrorString)>(sym("nvrtcGetErrorString"));
    g_nvrtc.Version           = reinterpret_cast<decltype(g_nvrtc.Version)>(sym("nvrtcVersion"));
    if (!ok) why = "libnvrtc lacks a required entry point";
    return ok;
}

bool load_driver(std::string& why) {
    bool ok  = true;
    auto get = [&](char const* n) -> void* {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t const e = cudaGetDriverEntryPoint(n, &fn, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || fn == nullptr) {
            cudaGetLastError();
            ok = false;
            if (why.empty()) why = std::string("the CUDA driver does not provide ") + n;
            return nullptr;
        }
        return fn;
    };
    g_drv.LibraryLoadData    = reinterpret_cast<decltype(g_drv.LibraryLoadData)>(get("cuLibraryLoadData"));
    g_drv.LibraryGetKernel   = reinterpret_cast<decltype(g_drv.LibraryGetKernel)>(get("cuLibraryGetKernel"));
    g_drv.KernelSetAttribute = reinterpret_cast<decltype(g_drv.KernelSetAttribute)>(get("cuKernelSetAttribute"));
    g_drv.LaunchKernel       = reinterpret_cast<decltype(g_drv.LaunchKernel)>(get("cuLaunchKernel"));
    g_drv.GetErrorString     = reinterpret_cast<decltype(g_drv.GetErrorString)>(get("cuGetErrorString"));
    g_drv.CtxGetDevice       = reinterpret_cast<decltype(g_drv.CtxGetDevice)>(get("cuCtxGetDevice"));
    {
        // programmatic dependent launch is an optimisation: its absence is not an error
        bool const was_ok = ok;
        std::string const was_why = why;
        g_drv.LaunchKernelEx = reinterpret_cast<decltype(g_drv.LaunchKernelEx)>(get("cuLaunchKernelEx"));
        ok = was_ok, why = was_why;
    }
    return ok;
}

// called with g_mu held
bool available_locked(bool need_driver) {
    if (g_avail < 0) {
        char const* env = std::getenv("QSX_JIT");
        if (env != nullptr && std::strcmp(env, "0") == 0) {
            g_avail = 0, g_why = "disabled by QSX_JIT=0";
        } else if (!load_nvrtc(g_why)) {
            g_avail = 0;
        } else {
            g_avail = 1;
        }
    }
    if (g_avail == 1 && need_driver && g_drv.LaunchKernel == nullptr) {
        std::string why;
        if (!load_driver(why)) {
            g_avail = 0, g_why = why;
        }
    }
    return g_avail == 1;
}

void hash_program(const unsigned char* p, size_t n, int R, int threads, uint64_t& h0, uint64_t& h1) {
    // two independent 64-bit multiplicative hashes over 8-byte words (blobs are 16-byte multiples)
    uint64_t a = 0x9e3779b97f4a7c15ull ^ (uint64_t)R, b = 0xc2b2ae3d27d4eb4full ^ ((uint64_t)threads << 8);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        a = (a ^ w) * 0x100000001b3ull;
        a ^= a >> 29;
        b = (b + w) * 0xff51afd7ed558ccdull;
        b ^= b >> 32;
    }
    for (; i < n; ++i) {
        a = (a ^ p[i]) * 0x100000001b3ull;
        b = (b + p[i]) * 0xff51afd7ed558ccdull;
    }
    h0 = a ^ (uint64_t)n, h1 = b;
}

constexpr bool kPersistentDefault = false;

int min_ctas_for(int threads) {
    if (char const* e = std::getenv("QSX_JIT_MINCTAS")) {
        int const v = std::atoi(e);
        if (v >= 1 && v <= 16) return v;
    }
    // 2^R elements of 2 doubles = 64 registers at R = 4; 128 registers per thread (16 warps per SM) measured best on
    // C4: 96 (20 warps) spills into the op stream, +13 % time; 170 (12 warps) hides the barriers worse
    // (profiles/r2_tile_sweep_c4.jsonl)
    int const v = 65536 / (threads * 128);
    return v < 1 ? 1 : (v > 8 ? 8 : v);
}

// ---- on-disk cache of compiled kernels ------------------------------------------------------------------
// One file per (generated source, compiler): a second process on the same host -- the other ranks of a torchrun job, the
// next invocation of the CLI on the same circuit -- loads the cubin instead of compiling it again.  The key is a hash of
// the SOURCE TEXT (so any change of the code generator or of the launch bounds is a different file) plus the NVRTC
// version.  QSX_JIT_CACHE=<dir> moves it, QSX_JIT_CACHE=off disables it.  Files are written to a temporary name and
// renamed, so a reader never sees a partial file.
std::string const& cache_dir() {
    // (never destroyed: its destructor would run at exit BEFORE the pool's exit hook, under the feet of the workers)
    static std::string const& dir = *new std::string([] {
        std::string d;
        if (char const* e = std::getenv("QSX_JIT_CACHE")) {
            d = e;
            if (d == "off" || d == "0") return std::string();
        } else if (char const* x = std::getenv("XDG_CACHE_HOME")) {
            d = std::string(x) + "/qsx-jit";
        } else {
            d = "/tmp/qsx-jit-" + std::to_string((unsigned)getuid());
        }
        ::mkdir(d.c_str(), 0700);
        struct stat st {};
        if (::stat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) return std::string();
        return d;
    }());
    return dir;
}

std::string cache_path(Kernel const* k) {
    if (cache_dir().empty()) return {};
    uint64_t h0, h1;
    hash_program(reinterpret_cast<const unsigned char*>(k->source.data()), k->source.size(), 0, 0, h0, h1);
    int major = 0, minor = 0;
    if (g_nvrtc.Version != nullptr) g_nvrtc.Version(&major, &minor);
    char nm[160];
    std::snprintf(nm, sizeof(nm), "/%016llx%016llx-sm100a-nvrtc%d.%d.cubin", (unsigned long long)h0, (unsigned long long)h1, major, minor);
    return cache_dir() + nm;
}

bool read_cached(std::string const& path, std::vector<char>& out) {
    if (path.empty()) return false;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (f == nullptr) return false;
    std::fseek(f, 0, SEEK_END);
    long const n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    bool ok = n > 64;
    if (ok) {
        out.resize((size_t)n);
        ok = std::fread(out.data(), 1, (size_t)n, f) == (size_t)n && std::memcmp(out.data(), "\x7f" "ELF", 4) == 0;
    }
    std::fclose(f);
    if (!ok) out.clear();
    return ok;
}

void write_cached(std::string const& path, std::vector<char> const& data) {
    if (path.empty()) return;
    std::string const tmp = path + ".tmp." + std::to_string((long)getpid()) + "." + std::to_string((unsigned long)(uintptr_t)&data);
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (f == nullptr) return;
    bool const ok = std::fwrite(data.data(), 1, data.size(), f) == data.size();
    std::fclose(f);
    if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}

void compile_one(Kernel* k) {
    auto const t0 = std::chrono::steady_clock::now();
    nvrtcProgram prog = nullptr;
    std::string err;
    bool ok = false, from_disk = false;
    bool generated = generate_sweep_source(k->program.data(), k->program.size(), k->R, k->name, k->threads, k->min_ctas, k->persistent,
                                           k->source, err);
    std::vector<unsigned char>().swap(k->program);
    std::string const cpath = generated ? cache_path(k) : std::string();
    if (generated && read_cached(cpath, k->cubin)) ok = from_disk = true;
    if (generated && !ok) do {
        nvrtcResult r = g_nvrtc.CreateProgram(&prog, k->source.c_str(), (k->name + ".cu").c_str(), 0, nullptr, nullptr);
        if (r != NVRTC_SUCCESS) {
            err = std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r);
            break;
        }
        // --fmad=false: the generated source spells every fma and every product out like the interpreter's handlers do
        char const* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-lineinfo"};
        r = g_nvrtc.CompileProgram(prog, 4, opts);
        if (r != NVRTC_SUCCESS) {
            size_t n = 0;
            g_nvrtc.GetProgramLogSize(prog, &n);
            std::string log(n, '\0');
            if (n > 0) g_nvrtc.GetProgramLog(prog, log.data());
            err = std::string("nvrtcCompileProgram: ") + g_nvrtc.GetErrorString(r) + "\n" + log.substr(0, 2000);
            break;
        }
        size_t n = 0;
        r = g_nvrtc.GetCUBINSize(prog, &n);
        if (r != NVRTC_SUCCESS || n == 0) {
            err = "nvrtcGetCUBINSize failed";
            break;
        }
        k->cubin.resize(n);
        r = g_nvrtc.GetCUBIN(prog, k->cubin.data());
        if (r != NVRTC_SUCCESS) {
            err = "nvrtcGetCUBIN failed";
            break;
        }
        ok = true;
    } while (false);
    if (prog != nullptr) g_nvrtc.DestroyProgram(&prog);
    if (ok && !from_disk) write_cached(cpath, k->cubin);
    if (char const* dir = std::getenv("QSX_JIT_DUMP")) {  // development aid: keep the generated source and the cubin
        std::string const base = std::string(dir) + "/" + k->name;
        if (FILE* f = std::fopen((base + ".cu").c_str(), "w")) {
            std::fwrite(k->source.data(), 1, k->source.size(), f);
            std::fclose(f);
        }
        if (ok)
            if (FILE* f = std::fopen((base + ".cubin").c_str(), "wb")) {
                std::fwrite(k->cubin.data(), 1, k->cubin.size(), f);
                std::fclose(f);
            }
    }
    double const dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    {
        std::lock_guard<std::mutex> lk(g_mu);
        // exit handlers run in reverse order of registration, and the compiler registers destructors of its own lazily
        // created state DURING the first compilations: the pool's hook is registered again behind them (bounded)
        if (!from_disk && g_hook_registrations < 64) {
            ++g_hook_registrations;
            std::atexit(stop_pool_at_exit);
        }
        g_stats.compile_seconds += dt;
        if (ok) {
            if (from_disk)
                ++g_stats.disk_hits;
            else
                ++g_stats.compiled;
        } else {
            ++g_stats.failed;
            g_last_error = err;
            if (std::getenv("QSX_JIT_VERBOSE") != nullptr) std::fprintf(stderr, "[qsx jit] %s: %s\n", k->name.c_str(), err.c_str());
        }
        std::string().swap(k->source);  // not needed any more
        k->state.store(ok ? kReady : kFailed, std::memory_order_release);
    }
    g_cv_done.notify_all();
}

void worker_main() {
    // the compilers fill every core; the thread that launches the sweeps (and the driver's own threads) must not queue
    // behind them: a Linux thread has its own nice value
    setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 10);
    for (;;) {
        Kernel* k = nullptr;
        {
            std::unique_lock<std::mutex> lk(g_mu);
            g_cv_work.wait(lk, [] { return g_stop || !g_queue.empty(); });
            if (g_stop) return;
            k = g_queue.front();
            g_queue.pop_front();
            k->state.store(kCompiling, std::memory_order_relaxed);
            ++g_active;
        }
        compile_one(k);
        {
            std::lock_guard<std::mutex> lk(g_mu);
            --g_active;
        }
        g_cv_done.notify_all();
    }
}

// A process must not run its static destructors (NVRTC has its own) while a compiler thread is still inside
// nvrtcCompileProgram: `quit` right after converting a 14-qubit circuit crashed exactly there.  At exit the queue is
// dropped (those sweeps were being interpreted anyway) and the compilations in flight are waited for (< 1 s).
void stop_pool_at_exit() {
    if (std::getenv("QSX_JIT_VERBOSE") != nullptr) std::fprintf(stderr, "[qsx jit] exit hook: waiting for %d compilation(s) in flight\n", g_active);
    shutdown();
    if (std::getenv("QSX_JIT_VERBOSE") != nullptr) std::fprintf(stderr, "[qsx jit] exit hook: done\n");
}

// called with g_mu held
void start_pool_locked() {
    if (g_pool_started) return;
    g_pool_started = true;
    int n = (int)std::thread::hardware_concurrency();
    if (char const* e = std::getenv("QSX_JIT_THREADS")) n = std::atoi(e);
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    g_stats.workers = n;
    std::atexit(stop_pool_at_exit);
    for (int i = 0; i < n; ++i) std::thread(worker_main).detach();  // daemon threads: they only ever wait on the queue
}

std::string cu_error(CUresult r) {
    const char* s = nullptr;
    if (g_drv.GetErrorString != nullptr && g_drv.GetErrorString(r, &s) == CUDA_SUCCESS && s != nullptr) return s;
    return "CUDA driver error " + std::to_string((int)r);
}

}  // namespace

bool available(std::string* why) {
    std::lock_guard<std::mutex> lk(g_mu);
    bool const ok = available_locked(false);
    if (!ok && why != nullptr) *why = g_why;
    return ok;
}

// Persistent sweeps (jit_codegen.cpp): QSX_JIT_PERSIST=0/1 overrides the default.
bool persistent_mode() {
    static bool const on = [] {
        char const* e = std::getenv("QSX_JIT_PERSIST");
        return e != nullptr ? std::atoi(e) != 0 : kPersistentDefault;
    }();
    return on;
}

Kernel* request(const unsigned char* blob, size_t size, int R, int threads) {
    uint64_t h0, h1;
    bool const persistent = persistent_mode();
    hash_program(blob, size, R, threads | (persistent ? 1 << 20 : 0), h0, h1);
    std::lock_guard<std::mutex> lk(g_mu);
    if (!available_locked(false)) return nullptr;
    ++g_stats.requested;
    auto it = g_cache.find({h0, h1});
    if (it != g_cache.end()) {
        ++g_stats.hits;
        return it->second.get();
    }
    auto k     = std::make_unique<Kernel>();
    k->h0 = h0, k->h1 = h1, k->R = R, k->threads = threads, k->min_ctas = min_ctas_for(threads), k->persistent = persistent;
    char nm[64];
    std::snprintf(nm, sizeof(nm), "qsx_sweep_%016llx%016llx", (unsigned long long)h0, (unsigned long long)h1);
    k->name = nm;
    if (g_stop) {
        k->state.store(kFailed);
    } else {
        // (the source is generated by the worker too: ~1 ms per sweep, which for the ~140 sweeps of a 15-qubit plan would
        //  otherwise sit on the caller's thread in front of the first launch)
        k->program.assign(blob, blob + size);
        start_pool_locked();
        g_queue.push_back(k.get());
        g_cv_work.notify_one();
    }
    Kernel* const raw = k.get();
    g_cache.emplace(std::make_pair(h0, h1), std::move(k));
    return raw;
}

unsigned cta_threads(Kernel* k) { return k != nullptr ? (unsigned)k->threads : 0u; }
bool ready(Kernel* k) { return k != nullptr && k->state.load(std::memory_order_acquire) == kReady && !k->load_failed; }
bool failed(Kernel* k) { return k == nullptr || k->state.load(std::memory_order_acquire) == kFailed || k->load_failed; }

bool wait(Kernel* k) {
    if (k == nullptr) return false;
    std::unique_lock<std::mutex> lk(g_mu);
    g_cv_done.wait(lk, [&] {
        int const s = k->state.load(std::memory_order_acquire);
        return s == kReady || s == kFailed;
    });
    return k->state.load() == kReady && !k->load_failed;
}

cudaError_t launch(Kernel* k, double2* u, int log2_ncols, int log2_ncb, uint32_t flags, uint32_t col0, unsigned grid, unsigned threads,
                   size_t smem, cudaStream_t stream) {
    if (!ready(k)) return cudaErrorNotReady;
    int dev = 0;
    if (cudaError_t const e = cudaGetDevice(&dev); e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lk(k->load_mu);
        if (k->kern == nullptr && !k->load_failed) {
            bool drv_ok;
            {
                std::lock_guard<std::mutex> g(g_mu);
                drv_ok = available_locked(true);
            }
            CUresult r = drv_ok ? g_drv.LibraryLoadData(&k->lib, k->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) : CUDA_ERROR_NOT_SUPPORTED;
            if (r == CUDA_SUCCESS) r = g_drv.LibraryGetKernel(&k->kern, k->lib, k->name.c_str());
            if (r != CUDA_SUCCESS) {
                k->load_failed = true;
                std::lock_guard<std::mutex> g(g_mu);
                g_last_error = "loading " + k->name + ": " + cu_error(r);
                ++g_stats.failed;
            } else {
                std::vector<char>().swap(k->cubin);  // the driver keeps its own copy
            }
        }
        if (k->load_failed) return cudaErrorNotReady;
        if (k->persistent) smem = ((size_t)16 << k->R) * (size_t)k->threads;
        if (dev >= 0 && dev < 64 && !k->attr_set[dev] && smem > 48 * 1024) {
            CUdevice cudev = 0;
            CUresult r     = g_drv.CtxGetDevice(&cudev);
            if (r == CUDA_SUCCESS) r = g_drv.KernelSetAttribute(CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 200 * 1024, k->kern, cudev);
            if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
            k->attr_set[dev] = true;
        }
    }
    // L2 prefetch distance in CTAs.  OFF by default: measured on C4 (profiles/r2_prefetch_sweep_c4.jsonl) every distance
    // from 1x to 8x the resident CTAs made the sweeps 14-40 % SLOWER -- the sweeps already run at the bandwidth this access
    // pattern (128-byte rows, 2^n x 16 B apart) gets from HBM, and the prefetches only add requests.  QSX_JIT_PREFETCH=<CTAs>
    // keeps the experiment reproducible.
    static uint32_t const pf_env = [] {
        char const* e = std::getenv("QSX_JIT_PREFETCH");
        return e != nullptr ? (uint32_t)std::max(0, std::atoi(e)) : 0u;
    }();
    uint32_t pf = pf_env;
    if (k->persistent) {
        static int sms[64] = {};
        if (dev >= 0 && dev < 64 && sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        unsigned const wave = (unsigned)((dev >= 0 && dev < 64 && sms[dev] > 0 ? sms[dev] : 148) * k->min_ctas);
        pf                  = grid;  // tile count
        grid                = grid < wave ? grid : wave;
    }
    uint32_t* ctr = nullptr;
    if (k->persistent) {  // the tile counter of the dynamic schedule, zeroed in stream order before every launch
        static uint32_t* ctrs[64] = {};
        if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
        if (ctrs[dev] == nullptr)
            if (cudaError_t const e = cudaMalloc(&ctrs[dev], 256); e != cudaSuccess) return e;
        ctr = ctrs[dev];
        if (cudaError_t const e = cudaMemsetAsync(ctr, 0, sizeof(uint32_t), stream); e != cudaSuccess) return e;
    }
    void* args[] = {&u, &log2_ncols, &log2_ncb, &flags, &col0, &pf, &ctr};
    // Programmatic dependent launch: a generated sweep signals `launch_dependents` when it starts and executes
    // `griddepcontrol.wait` before its first global access (jit_codegen.cpp), so the NEXT sweep's CTAs are scheduled and run
    // their index prologue while the last wave of this one drains; the data dependence itself is untouched (the wait returns
    // when the previous grid has completed and its writes are visible).  Worth ~1-2 us per launch: nothing for a 6 ms sweep,
    // 10 % for the 15 us sweeps of a 32 MiB shard.  QSX_JIT_PDL=0 launches the plain way.
    static bool const pdl = [] {
        char const* e = std::getenv("QSX_JIT_PDL");
        return e == nullptr || std::atoi(e) != 0;
    }();
    CUresult r;
    if (pdl && g_drv.LaunchKernelEx != nullptr) {
        CUlaunchAttribute attr{};
        attr.id                                         = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
        attr.value.programmaticStreamSerializationAllowed = 1;
        CUlaunchConfig cfg{};
        cfg.gridDimX = grid, cfg.gridDimY = 1, cfg.gridDimZ = 1;
        cfg.blockDimX = threads, cfg.blockDimY = 1, cfg.blockDimZ = 1;
        cfg.sharedMemBytes = (unsigned)smem;
        cfg.hStream        = stream;
        cfg.attrs          = &attr;
        cfg.numAttrs       = 1;
        r = g_drv.LaunchKernelEx(&cfg, reinterpret_cast<CUfunction>(k->kern), args, nullptr);
    } else {
        r = g_drv.LaunchKernel(reinterpret_cast<CUfunction>(k->kern), grid, 1, 1, threads, 1, 1, (unsigned)smem, stream, args, nullptr);
    }
    if (r != CUDA_SUCCESS) {
        std::lock_guard<std::mutex> g(g_mu);
        g_last_error = "launching " + k->name + ": " + cu_error(r);
        return cudaErrorLaunchFailure;
    }
    {
        std::lock_guard<std::mutex> g(g_mu);
        ++g_stats.launches;
    }
    return cudaSuccess;
}

void shutdown() {
    std::unique_lock<std::mutex> lk(g_mu);
    if (!g_pool_started) return;
    g_stop = true;
    for (Kernel* k : g_queue) k->state.store(kFailed, std::memory_order_release);
    g_queue.clear();
    g_cv_work.notify_all();
    g_cv_done.notify_all();
    g_cv_done.wait(lk, [] { return g_active == 0; });
    // the workers are gone; a later request starts a new pool
    g_pool_started = false;
    g_stop         = false;
}

Stats stats() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_stats;
}

std::string last_compile_error() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_last_error;
}

}  // namespace qsx::jit
