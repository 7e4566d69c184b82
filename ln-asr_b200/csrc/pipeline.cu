// pipeline.cu -- one C call for the whole decode path of a batch: PCM -> MFCC -> GMM emissions -> Viterbi -> word.
//
// Chains the entry points of this library that replace MFCC.calc_mfcc (/root/reference/lnasr/mfcc.py:108-175),
// GMMHMM._map_b (gmmhmm.py:46-67), HMM._viterbi / decode (hmm.py:133-172, 276-283) and the isolated-word decision
// (argmax over models; Recognizer.recognize is `pass` in the reference, recognizer.py:46-48) on ONE
// caller-provided workspace and ONE stream: 7-8 kernel launches, no allocation, no synchronisation.  It exists for the
// host-audio path, where a batch is cut into groups whose copies and compute overlap: the host cost per group drops from
// ~15 Python-level calls to one.
#include "common.cuh"

using namespace lnasr;

static inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

struct PipeLayout {
    size_t feats, mfcc_ws, mfcc_ws_bytes, b, gmm_ws, gmm_ws_bytes, vit_ws, vit_ws_bytes, scores, total;
};

static int pipe_layout(int64_t total_frames, int32_t n_utt, const lnasr_recognizer* r, int want_paths, PipeLayout* L) {
    const int64_t ncol = (int64_t)r->bank->n_models * r->bank->N;
    const int ncf = 3 * (r->mfcc->n_ceps + 1);
    size_t o = 0;
    L->feats = o; o += up256((size_t)total_frames * ncf * sizeof(float));
    size_t mw = 0;
    int rc = lnasr_mfcc_workspace_bytes(total_frames, n_utt, &mw);
    if (rc) return rc;
    L->mfcc_ws = o; L->mfcc_ws_bytes = mw; o += up256(mw);
    L->b = o; o += up256((size_t)total_frames * ncol * (r->gmm_kind >= 1 ? sizeof(float) : sizeof(double)));
    size_t gw = 0;
    if (r->gmm_kind >= 1) {
        rc = lnasr_gmm_tc_workspace_bytes(total_frames, &gw);
        if (rc) return rc;
    }
    L->gmm_ws = o; L->gmm_ws_bytes = gw; o += up256(gw);
    const size_t vw = want_paths ? lnasr_viterbi_workspace_bytes(total_frames, n_utt, r->bank->n_models, r->bank->N) : 0;
    L->vit_ws = o; L->vit_ws_bytes = vw; o += up256(vw);
    L->scores = o; o += up256((size_t)n_utt * r->bank->n_models * sizeof(double));
    L->total = o;
    return LNASR_OK;
}

static int check_rec(const lnasr_recognizer* r, const char* who) {
    if (!r || !r->mfcc || !r->bank || !r->gmm_packed) { set_error("%s: null pointer in lnasr_recognizer", who); return LNASR_EINVAL; }
    if (r->gmm_kind < 0 || r->gmm_kind > 2) { set_error("%s: gmm_kind must be 0 (exact), 1 (tc) or 2 (tc16)", who); return LNASR_EINVAL; }
    if (r->dim != 3 * (r->mfcc->n_ceps + 1)) { set_error("%s: model dim %d != %d MFCC features", who, r->dim, 3 * (r->mfcc->n_ceps + 1)); return LNASR_EINVAL; }
    return LNASR_OK;
}

extern "C" int lnasr_recognize_workspace_bytes(int64_t total_frames, int32_t n_utt, const lnasr_recognizer* rec,
                                               int want_paths, size_t* bytes) {
    int rc = check_rec(rec, "recognize_workspace_bytes");
    if (rc) return rc;
    LNASR_CHECK_ARG(bytes && total_frames >= 0 && n_utt >= 0, "recognize_workspace_bytes: bad arguments");
    PipeLayout L;
    rc = pipe_layout(total_frames, n_utt, rec, want_paths, &L);
    if (rc) return rc;
    *bytes = L.total;
    return LNASR_OK;
}

extern "C" int lnasr_recognize_batch(const void* pcm, int pcm_dtype, const int64_t* sample_offsets,
                                     const int64_t* frame_offsets, int32_t n_utt, int64_t total_frames,
                                     const lnasr_recognizer* rec, int32_t* words_out, double* best_out,
                                     uint32_t* paths_out, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_rec(rec, "recognize_batch");
    if (rc) return rc;
    LNASR_CHECK_ARG(pcm && sample_offsets && frame_offsets && words_out && workspace, "recognize_batch: null pointer");
    if (n_utt == 0 || total_frames == 0) return LNASR_OK;
    PipeLayout L;
    rc = pipe_layout(total_frames, n_utt, rec, paths_out != nullptr, &L);
    if (rc) return rc;
    if (workspace_bytes < L.total) {
        set_error("recognize_batch: workspace %zu < %zu bytes", workspace_bytes, L.total);
        return LNASR_EWORKSPACE;
    }
    char* ws = (char*)workspace;
    const int ncf = 3 * (rec->mfcc->n_ceps + 1);
    const int64_t ncol = (int64_t)rec->bank->n_models * rec->bank->N;
    float* feats = (float*)(ws + L.feats);
    rc = lnasr_mfcc_batch(pcm, pcm_dtype, sample_offsets, frame_offsets, n_utt, total_frames, rec->mfcc, feats, ncf, nullptr,
                          nullptr, ws + L.mfcc_ws, L.mfcc_ws_bytes, stream);
    if (rc) return rc;
    void* b = ws + L.b;
    if (rec->gmm_kind == 1)
        rc = lnasr_gmm_loglik_tc(feats, ncf, total_frames, rec->gmm_packed, (int32_t)ncol, rec->n_mix, rec->dim, (float*)b, ncol,
                                 ws + L.gmm_ws, L.gmm_ws_bytes, stream);
    else if (rec->gmm_kind == 2)
        rc = lnasr_gmm_loglik_tc16(feats, ncf, total_frames, rec->gmm_packed, (int32_t)ncol, rec->n_mix, rec->dim, (float*)b, ncol,
                                   ws + L.gmm_ws, L.gmm_ws_bytes, stream);
    else
        rc = lnasr_gmm_loglik_exact(feats, LNASR_DT_F32, ncf, total_frames, (const double*)rec->gmm_packed, (int32_t)ncol, rec->n_mix,
                                    rec->dim, b, LNASR_DT_F64, ncol, nullptr, stream);
    if (rc) return rc;
    double* scores = (double*)(ws + L.scores);
    rc = lnasr_viterbi_batch(b, rec->gmm_kind >= 1 ? LNASR_DT_F32 : LNASR_DT_F64, ncol, frame_offsets, n_utt, total_frames, rec->bank,
                             paths_out, scores, nullptr, nullptr, paths_out ? ws + L.vit_ws : nullptr, L.vit_ws_bytes, stream);
    if (rc) return rc;
    return lnasr_argmax_rows(scores, n_utt, rec->bank->n_models, words_out, best_out, stream);
}
