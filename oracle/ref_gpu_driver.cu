/* TEST / BASELINE INFRASTRUCTURE -- headless driver around the UNMODIFIED reference
 * translation unit.  oracle/make_ref.py compiles this file with
 *     nvcc -std=c++17 -O3 -arch=sm_100 -Ioracle/stubs -D_USE_MATH_DEFINES
 *          -Dmain=ref_main -DREF_KERNEL_CU="/root/reference/raygpu/kernel.cu"
 * so the #include below pulls in the reference where it lies; nothing of it is
 * copied into this repo.  The result (oracle/_ref/libdogeray_ref_gpu.so) is the
 * "reference kernel.cu rebuilt for sm_100" baseline of BASELINE.md section 2.
 *
 * Two timings are exposed:
 *   refgpu_frame        one CudaStarter() call exactly as the product runs it
 *                       (malloc + full scene upload + launch + sync + download + free,
 *                       kernel.cu:2562-2669), wall clock
 *   refgpu_kernel_only  the same Kernel<<<(W/div/8,H/div/8),(8,8)>>> launch
 *                       (kernel.cu:2634-2640) from buffers kept resident, CUDA events
 */
#include <unistd.h>
#include <limits.h>
#include <cuda_runtime.h>
#ifdef REF_COUNT_RAYS
/* The ray-counting variant (libdogeray_ref_gpu_count.so): make_ref.py pipes kernel.cu through sed, which puts a call
 * to refgpu_count_ray() in front of the single hit() call of raycolor (kernel.cu:800), into a temporary file outside
 * the repo.  It exists to let the reference count ITS OWN rays; it is never the build that is timed. */
__device__ unsigned long long refgpu_ray_counter;
__device__ __forceinline__ void refgpu_count_ray()
{
    const unsigned m = __activemask();
    const unsigned lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31u;
    if (lane == (unsigned)(__ffs(m) - 1)) atomicAdd(&refgpu_ray_counter, (unsigned long long)__popc(m));
}
#endif
#include REF_KERNEL_CU

static singleobject* g_objs = nullptr;
static bvh* g_nodes = nullptr;
static cudaTextureObject_t* g_tex = nullptr;
static std::string* g_texpaths = nullptr;

/* resident copies for the kernel-only timing */
static float* d_settings = nullptr;
static int3* d_out = nullptr;
static bvh* d_nodes = nullptr;
static singleobject* d_objs = nullptr;
static cudaTextureObject_t* d_tex = nullptr;

__global__ void refgpu_ids_kernel(const float* o3, const float* d3, int n, float* t_out, int* id_out, bvh* nodes, singleobject* objs)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 r = hit(make_float3(o3[3*i], o3[3*i+1], o3[3*i+2]), make_float3(d3[3*i], d3[3*i+1], d3[3*i+2]), nodes, objs);
    t_out[i] = r.x;
    id_out[i] = (r.x > 0.0) ? (int)r.y : -1;
}

static void free_resident()
{
    cudaFree(d_settings); cudaFree(d_out); cudaFree(d_nodes); cudaFree(d_objs); cudaFree(d_tex);
    d_settings = nullptr; d_out = nullptr; d_nodes = nullptr; d_objs = nullptr; d_tex = nullptr;
}

extern "C" {

void refgpu_free()
{
    free_resident();
    delete[] g_objs; g_objs = nullptr;
    delete[] g_nodes; g_nodes = nullptr;
    delete[] g_tex; g_tex = nullptr;
    delete[] g_texpaths; g_texpaths = nullptr;
}

/* main()'s start-up sequence, kernel.cu:2055-2103, with the process CWD pointed at tex_dir
 * for the duration of the texture scan (the reference scans its CWD) */
int refgpu_load(const char* rts_path, const char* tex_dir)
{
    refgpu_free();
    SCREEN_WIDTH = 1280; SCREEN_HEIGHT = 720;
    objnum = 10000; bvhnum = objnum * 2; texnum = 1; iter = 0; backtex = -1;
    campos = make_float3(0, 0, 2); look = make_float3(0, 0, 0);
    aperturee = 0.01f; focus_diste = 3; actualbvhnum = 0; max_depthh = 50; samples_per_pixell = 1; fovv = 45;
    nbackgroundintensity[0] = 1; nanum[0] = 0;

    char abs_rts[PATH_MAX];
    if (!realpath(rts_path, abs_rts)) return -1;
    char old_cwd[PATH_MAX];
    if (!getcwd(old_cwd, sizeof old_cwd)) return -1;
    if (tex_dir && tex_dir[0] && chdir(tex_dir) != 0) return -2;

    objnum = getnum(abs_rts);
    if (objnum <= 0) { if (chdir(old_cwd)) {} return -1; }
    g_objs = new singleobject[objnum];
    texnum = getppmnum();
    g_texpaths = new std::string[texnum > 0 ? texnum : 1];
    getppmpaths(g_texpaths);
    read(abs_rts, g_objs, g_texpaths);
    bvhnum = nanum[0] * 2;
    g_nodes = new bvh[bvhnum];
    g_tex = new cudaTextureObject_t[texnum > 0 ? texnum : 1];
    readtextures(g_tex, g_texpaths);
    if (nanum[0] - 1 >= 2) build_bvh(g_nodes, g_objs);     /* see ref_host_api.inc: fewer than two objects never terminates */
    else return -4;
    nbvhnumnum[0] = bvhnum;
    cudaMemcpyToSymbol(anum, &nanum[0], sizeof(int), 0, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(dbvhnumnum, &nbvhnumnum[0], sizeof(int), 0, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(edebugnum, &debugnum[0], sizeof(int), 0, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(backgroundintensity, &nbackgroundintensity[0], sizeof(float), 0, cudaMemcpyHostToDevice);
    if (chdir(old_cwd)) {}
    return cudaGetLastError() == cudaSuccess ? objnum : -3;
}

void refgpu_get_settings(float* out)
{
    out[0] = campos.x; out[1] = campos.y; out[2] = campos.z; out[3] = aperturee;
    out[4] = look.x; out[5] = look.y; out[6] = look.z; out[7] = focus_diste;
    out[8] = (float)fovv; out[9] = (float)max_depthh; out[10] = (float)samples_per_pixell;
    out[11] = nbackgroundintensity[0]; out[12] = (float)backtex;
    out[13] = (float)SCREEN_WIDTH; out[14] = (float)SCREEN_HEIGHT; out[15] = 0;
}

void refgpu_set_settings(const float* in)
{
    campos = make_float3(in[0], in[1], in[2]); aperturee = in[3];
    look = make_float3(in[4], in[5], in[6]); focus_diste = in[7];
    fovv = (int)in[8]; max_depthh = (int)in[9]; samples_per_pixell = (int)in[10];
    nbackgroundintensity[0] = in[11]; backtex = (int)in[12];
    SCREEN_WIDTH = (int)in[13]; SCREEN_HEIGHT = (int)in[14];
    cudaMemcpyToSymbol(backgroundintensity, &nbackgroundintensity[0], sizeof(float), 0, cudaMemcpyHostToDevice);
    free_resident();
}

int refgpu_num_objects() { return nanum[0] - 1; }

/* rays the reference's own raycolor traced since the last reset (counting variant only; -1 otherwise) */
long long refgpu_rays(int reset)
{
#ifdef REF_COUNT_RAYS
    unsigned long long v = 0, zero = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&v, refgpu_ray_counter, sizeof v);
    if (reset) cudaMemcpyToSymbol(refgpu_ray_counter, &zero, sizeof zero);
    return (long long)v;
#else
    (void)reset;
    return -1;
#endif
}

/* one CudaStarter() call; out = W*H*3 ints indexed x*H+y; returns the cudaError_t it returned */
int refgpu_frame(int* out, int divisor, double* wall_ms)
{
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    cudaError_t e = CudaStarter((int3*)out, g_nodes, g_objs, g_tex, divisor);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (wall_ms) *wall_ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
    return (int)e;
}

/* `iters` launches of the reference kernel from resident buffers; returns mean ms per launch (<0 on error) */
float refgpu_kernel_only(int* out, int divisor, int iters)
{
    float settings[13] = { campos.x, campos.y, campos.z, look.x, look.y, look.z, aperturee, focus_diste,
                           (float)fovv, (float)max_depthh, (float)samples_per_pixell, (float)divisor, (float)backtex };
    size_t size = (size_t)SCREEN_WIDTH * SCREEN_HEIGHT;
    if (!d_out) {
        cudaMalloc(&d_out, size * sizeof(int3));
        cudaMalloc(&d_settings, 13 * sizeof(float));
        cudaMalloc(&d_nodes, (size_t)bvhnum * sizeof(bvh));
        cudaMalloc(&d_objs, (size_t)objnum * sizeof(singleobject));
        cudaMalloc(&d_tex, (size_t)(texnum > 0 ? texnum : 1) * sizeof(cudaTextureObject_t));
        cudaMemcpy(d_nodes, g_nodes, (size_t)bvhnum * sizeof(bvh), cudaMemcpyHostToDevice);
        cudaMemcpy(d_objs, g_objs, (size_t)objnum * sizeof(singleobject), cudaMemcpyHostToDevice);
        cudaMemcpy(d_tex, g_tex, (size_t)texnum * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice);
    }
    cudaMemcpy(d_settings, settings, sizeof settings, cudaMemcpyHostToDevice);
    dim3 threadsPerBlock(8, 8);
    dim3 numBlocks(SCREEN_WIDTH / divisor / threadsPerBlock.x, SCREEN_HEIGHT / divisor / threadsPerBlock.y);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i)
        Kernel<<<numBlocks, threadsPerBlock>>>(d_out, d_settings, d_nodes, d_objs, d_tex, SCREEN_WIDTH, SCREEN_HEIGHT);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) return -1.0f;
    if (out) cudaMemcpy(out, d_out, size * sizeof(int3), cudaMemcpyDeviceToHost);
    return ms / (iters > 0 ? iters : 1);
}

/* closest-hit ids through the reference's device hit() as nvcc compiles it (default -fmad=true) */
int refgpu_ids(const float* o3, const float* d3, int n, float* t_out, int* id_out)
{
    float *do3, *dd3, *dt; int* did;
    bvh* dn; singleobject* dob;
    cudaMalloc(&do3, (size_t)n * 12); cudaMalloc(&dd3, (size_t)n * 12); cudaMalloc(&dt, (size_t)n * 4); cudaMalloc(&did, (size_t)n * 4);
    cudaMalloc(&dn, (size_t)bvhnum * sizeof(bvh)); cudaMalloc(&dob, (size_t)objnum * sizeof(singleobject));
    cudaMemcpy(do3, o3, (size_t)n * 12, cudaMemcpyHostToDevice);
    cudaMemcpy(dd3, d3, (size_t)n * 12, cudaMemcpyHostToDevice);
    cudaMemcpy(dn, g_nodes, (size_t)bvhnum * sizeof(bvh), cudaMemcpyHostToDevice);
    cudaMemcpy(dob, g_objs, (size_t)objnum * sizeof(singleobject), cudaMemcpyHostToDevice);
    refgpu_ids_kernel<<<(n + 127) / 128, 128>>>(do3, dd3, n, dt, did, dn, dob);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(t_out, dt, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(id_out, did, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(do3); cudaFree(dd3); cudaFree(dt); cudaFree(did); cudaFree(dn); cudaFree(dob);
    return (int)e;
}

} /* extern "C" */
