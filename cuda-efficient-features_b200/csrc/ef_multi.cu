// ef_multi.cu -- single-process multi-GPU driver of the hot path: frames are independent, so a batch is cut into
// contiguous blocks, one per device, and every device runs the unmodified single-GPU pipeline (ef_detect_and_compute_host_batch)
// on its block from its own host thread and stream.  No cross-GPU exchange on the data path: the images go host -> owning
// device only, results come back device -> host.  (The per-process form of the same sharding -- one rank per GPU under
// torchrun -- lives in efb200/sharding.py and bench.py.)
#include "ef_common.cuh"

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

struct ef_mg_handle {
    std::vector<ef_handle*> dev;      // one single-GPU handle per device
    std::vector<int> ordinal;
    std::vector<cudaStream_t> stream;
    int max_batch = 1;
    std::string err;
};

extern "C" {

int ef_mg_create(const ef_params* params, const int* devices, int ndev, ef_mg_handle** out)
{
    if (!params || !out || ndev < 1) return EF_ERR_BAD_ARG;
    *out = nullptr;
    ef_mg_handle* m = new ef_mg_handle();
    m->max_batch = std::max(1, params->max_batch);
    for (int i = 0; i < ndev; i++) {
        ef_params p = *params;
        p.device = devices ? devices[i] : i;
        ef_handle* h = nullptr;
        const int rc = ef_create(&p, &h);
        cudaStream_t s = nullptr;
        if (rc == EF_OK && (cudaSetDevice(p.device) != cudaSuccess || cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess)) {
            ef_destroy(h);
            h = nullptr;
        }
        if (!h) {
            for (size_t j = 0; j < m->dev.size(); j++) { cudaSetDevice(m->ordinal[j]); cudaStreamDestroy(m->stream[j]); ef_destroy(m->dev[j]); }
            delete m;
            return rc != EF_OK ? rc : EF_ERR_CUDA;
        }
        m->dev.push_back(h); m->ordinal.push_back(p.device); m->stream.push_back(s);
    }
    *out = m;
    return EF_OK;
}

void ef_mg_destroy(ef_mg_handle* m)
{
    if (!m) return;
    for (size_t j = 0; j < m->dev.size(); j++) { cudaSetDevice(m->ordinal[j]); cudaStreamDestroy(m->stream[j]); ef_destroy(m->dev[j]); }
    delete m;
}

int ef_mg_device_count(const ef_mg_handle* m) { return m ? (int)m->dev.size() : 0; }
const char* ef_mg_last_error_string(const ef_mg_handle* m) { return m ? m->err.c_str() : "null handle"; }

// Frames [begin, end) of device i: contiguous blocks that differ by at most one frame (same rule as efb200.sharding.shard_range).
void ef_mg_shard_range(int nframes, int i, int ndev, int* begin, int* end)
{
    const int base = nframes / ndev, extra = nframes % ndev;
    *begin = i * base + std::min(i, extra);
    *end = *begin + base + (i < extra ? 1 : 0);
}

int ef_mg_detect_and_compute_host_batch(ef_mg_handle* m, int nframes, const uint8_t* h_imgs, size_t img_stride, size_t pitch,
                                        int width, int height, float* h_kpts5, uint8_t* h_desc, int* h_counts)
{
    if (!m) return EF_ERR_BAD_ARG;
    if (nframes < 0 || (nframes > 0 && (!h_imgs || !h_kpts5 || !h_counts))) { m->err = "null host pointer"; return EF_ERR_BAD_ARG; }
    const int ndev = (int)m->dev.size();
    std::vector<int> rc(ndev, EF_OK);
    double nf = 0, db = 0;
    ef_get_param(m->dev[0], EF_PARAM_MAX_FEATURES, &nf);
    db = ef_descriptor_size(m->dev[0]);
    auto work = [&](int i) {
        int b, e;
        ef_mg_shard_range(nframes, i, ndev, &b, &e);
        for (int f = b; f < e && rc[i] == EF_OK; f += m->max_batch) {
            const int n = std::min(m->max_batch, e - f);
            rc[i] = ef_detect_and_compute_host_batch(m->dev[i], n, h_imgs + (size_t)f * img_stride, img_stride, pitch, width, height,
                                                     h_kpts5 + (size_t)f * EF_ROWS_COUNT * (size_t)nf,
                                                     h_desc ? h_desc + (size_t)f * (size_t)nf * (size_t)db : nullptr, h_counts + f, m->stream[i]);
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < ndev; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < ndev; i++)
        if (rc[i] != EF_OK) {
            m->err = "device " + std::to_string(m->ordinal[i]) + ": " + ef_last_error_string(m->dev[i]);
            return rc[i];
        }
    return EF_OK;
}

} // extern "C"
