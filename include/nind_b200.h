/* nind_b200.h — C ABI of the B200-native NIND denoiser hot path.
 *
 * Drop-in boundary for the tiled inference path of esq4/nind-denoise (paths relative to the
 * reference tree):
 *   - network forward        src/nind_denoise/networks/UtNet.py:97-109 (UtNet.forward)
 *                            src/nind_denoise/networks/ThirdPartyNets.py:154-169 (UNet.forward)
 *   - model construction     src/nind_denoise/nn_common.py:116-138 (Model.instantiate_model)
 *   - crop grid / gather     src/nind_denoise/denoise_image.py:88-174 (OneImageDS)
 *   - trim / seam / stitch   src/nind_denoise/denoise_image.py:204-213, 240-267
 *
 * Conventions: every function returns 0 on success or a negative NIND_E_* code and never throws;
 * nind_last_error() returns a human-readable description of the most recent failure on the
 * calling thread.  The caller owns every buffer it passes in; the library owns its packed
 * weights and activation arena.  A handle is bound to the CUDA device that was current when it was
 * created (every entry point switches to that device for the duration of the call; nind_net_device()
 * reports it); it is not thread-safe, different handles are independent — also on different devices of
 * one process.  All work is enqueued on the `stream` argument (a cudaStream_t passed as void*; NULL =
 * default stream) and is complete once that stream is synchronised — except the *_host entry point, which
 * synchronises itself.  Consecutive calls on one handle may use different streams: the scratch memory the
 * entry points share is ordered by an internal event, no synchronisation is needed in between.
 * There is no CPU fallback: without an sm_100 device every compute entry point fails.
 */
#ifndef NIND_B200_H
#define NIND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nind_net nind_net;

enum {
  NIND_OK = 0,
  NIND_E_INVALID = -1,     /* bad argument (illegal crop size, unknown activation, ...) */
  NIND_E_CUDA = -2,        /* CUDA runtime / driver error */
  NIND_E_WEIGHTS = -3,     /* missing or mis-shaped tensor in the state_dict */
  NIND_E_KERNEL = -4,      /* a kernel reported a pipeline time-out */
  NIND_E_UNSUPPORTED = -5  /* configuration not implemented by this build */
};

enum { NIND_ARCH_UTNET = 0, NIND_ARCH_UNET = 1 };
enum { NIND_ACT_PRELU = 0, NIND_ACT_ELU = 1, NIND_ACT_HARDSWISH = 2 };
enum { NIND_FWD_CLAMP01 = 1 };                           /* nind_net_forward_ex flags */
enum { NIND_PIX_U8 = 0, NIND_PIX_U16 = 1, NIND_PIX_F32 = 2 };  /* interleaved image pixel types */

/* One entry of a PyTorch state_dict: fp32, contiguous, host or device memory. */
typedef struct {
  const char* name;     /* e.g. "convs1.0.weight" (UtNet.py) / "down1.mpconv.1.conv.0.weight" (UNet) */
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} nind_tensor;

/* One row of the crop table (denoise_image.py:130-137,172-173). */
typedef struct {
  int32_t x0, y0;                       /* top-left of the cs x cs window in the image (may be < 0) */
  int32_t ud_x0, ud_y0, ud_x1, ud_y1;   /* usefuldim inside the crop                                */
  int32_t start_x, start_y;             /* usefulstart: position of the useful area in the image    */
} nind_crop;

/* Library / device information.  Returns 0 and fills *sm_major/minor if a CUDA device is usable. */
int nind_device_info(int* sm_major, int* sm_minor, int* sm_count);

/* Replaces: globals()[network](**parameters) + load_state_dict (nn_common.py:129-132).
 * `arch` NIND_ARCH_*, `funit` 64 (UtNet) — ignored by UNet as in the reference
 * (ThirdPartyNets.py:139-150) —, `activation` NIND_ACT_* (UtNet.py:17-26).
 * Packs the fp32 tensors into the bf16 tap-major layout the kernels use (BatchNorm folded). */
int nind_net_create(int arch, int funit, int activation, const nind_tensor* tensors, int n_tensors,
                    nind_net** out);
/* Re-pack after the caller changed its parameters (load_state_dict on an existing module). */
int nind_net_load(nind_net* net, const nind_tensor* tensors, int n_tensors);
void nind_net_destroy(nind_net* net);

/* Replaces: model(ybatch) (denoise_image.py:246).  in/out: device pointers, fp32 NCHW
 * [batch,3,h,w] contiguous.  UtNet: h and w must be 16a+56, a >= 3 (UtNet.py:6-7); UNet: multiples
 * of 16. */
int nind_net_forward(nind_net* net, const float* in_nchw, float* out_nchw, int batch, int h, int w,
                     void* stream);
/* Same with flags.  NIND_FWD_CLAMP01 replaces Generator.denoise_batch = model(x).clip(0, 1)
 * (nn_common.py:198-199, used by the validation / test passes of nn_train.py:51-93): the clamp is fused
 * into the epilogue of the 1x1 output head. */
int nind_net_forward_ex(nind_net* net, const float* in_nchw, float* out_nchw, int batch, int h, int w, int flags,
                        void* stream);
/* CUDA device ordinal the handle is bound to. */
int nind_net_device(nind_net* net, int* device);

/* Crop grid (OneImageDS.__init__/__getitem__).  Pass table = NULL to query the count. */
int nind_crop_table(int width, int height, int cs, int ucs, int ol, nind_crop* table, int* n_crops);

/* Replaces the main loop of denoise_image.py:240-267 for crops [crop_begin, crop_end) of a planar
 * fp32 [3,height,width] device image: gather (mirror pad) -> forward in batches of `batch` crops ->
 * trim -> seam halving -> overlap-add.  Writes the rows [*band_y0, *band_y1) those crops touch into
 * out_band (device, planar [3, band rows, width], at least nind_band_rows() rows); pixels inside
 * the band that belong to other crops' useful areas receive only this range's contributions
 * (add the bands of all ranges to obtain the image). */
int nind_tiled_denoise(nind_net* net, const float* img_chw, float* out_band, int height, int width,
                       int cs, int ucs, int ol, int crop_begin, int crop_end, int batch,
                       int* band_y0, int* band_y1, void* stream);
int nind_band_rows(int width, int height, int cs, int ucs, int ol, int crop_begin, int crop_end,
                   int* band_y0, int* band_y1);

/* Step-wise form of nind_tiled_denoise for callers that overlap communication with compute (the multi-GPU gather
 * of the stitched output): crops [step_begin, step_end) of the range [crop_begin, crop_end) go through one
 * forward, and the band rows that become final with them, [*rows_begin, *rows_end), are stitched into out_img
 * — a device image of FULL [3,height,width] layout.  Steps must be issued in raster order and cover the range
 * (nind_plan_steps gives the boundaries the host pipeline uses: bounds[0..n_steps]). */
int nind_tiled_denoise_step(nind_net* net, const float* img_chw, float* out_img, int height, int width, int cs,
                            int ucs, int ol, int crop_begin, int crop_end, int step_begin, int step_end,
                            int* rows_begin, int* rows_end, void* stream);
int nind_plan_steps(nind_net* net, int width, int height, int cs, int ucs, int ol, int crop_begin, int crop_end,
                    int batch, int* bounds, int max_bounds, int* n_steps);
/* dst[p*dst_plane + i] = src[p*src_plane + i] for p < planes, i < count: asynchronous device-to-device copies on
 * `stream` (copy engines, no SM).  dst may be peer memory opened with nind_peer_open: the bytes then go straight
 * over NVLink. */
int nind_copy_planes(float* dst, long long dst_plane, const float* src, long long src_plane, int planes,
                     long long count, void* stream);
/* Peer memory for the multi-GPU gather of the stitched output: nind_peer_alloc allocates device memory on the
 * current device and returns its 64-byte CUDA IPC handle; another process of the node maps it on ITS current
 * device with nind_peer_open (lazy peer access), after which copies and stores from that device reach it over
 * NVLink.  The reference has no counterpart (single device). */
int nind_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int nind_peer_open(const unsigned char* handle64, void** ptr);
int nind_peer_close(void* ptr);
int nind_peer_free(void* ptr);
/* dst[p*dst_plane + i] += src[p*src_plane + i] for p < planes, i < count (fp32 device pointers; vectorised when
 * count, the plane strides and the pointers are multiples of 4 floats): the partial sums another rank computed
 * for image rows are added to the rows' owner's, in rank = raster order. */
int nind_add_rows(float* dst, long long dst_plane, const float* src, long long src_plane, int planes, long long count,
                  void* stream);

/* The two geometry halves of the loop as stand-alone ops (bit-exact copies / fp32 adds):
 *   nind_gather_crops  = OneImageDS.__getitem__ for crops [crop_begin, crop_end)
 *                        (denoise_image.py:129-174) -> crops_out [n,3,cs,cs] fp32 device;
 *   nind_stitch_crops  = trim + make_seamless_edges + overlap-add (denoise_image.py:204-213,250-267)
 *                        of network outputs crops [n,3,cs,cs] into the row band of that range. */
int nind_gather_crops(nind_net* net, const float* img_chw, int height, int width, int cs, int ucs, int ol,
                      int crop_begin, int crop_end, float* crops_out, void* stream);
int nind_stitch_crops(const float* crops, int height, int width, int cs, int ucs, int ol, int crop_begin,
                      int crop_end, float* out_band, int* band_y0, int* band_y1, void* stream);

/* Whole image, host buffers (pageable or pinned): H2D, all crops, D2H, synchronised.  This is the
 * call the reference-facing plugin makes for one image on one GPU. */
int nind_tiled_denoise_host(nind_net* net, const float* img_chw_host, float* out_chw_host, int height,
                            int width, int cs, int ucs, int ol, int batch);

/* Throughput mode (a stream of images on one GPU, BASELINE config "batch of 64 x 24 MP"): enqueue without
 * waiting — image k+1's upload overlaps image k's compute and download (two device slots) — then
 * nind_host_sync() once.  Host buffers must stay valid (and should be pinned) until the sync returns. */
int nind_tiled_denoise_host_async(nind_net* net, const float* img_chw_host, float* out_chw_host, int height,
                                  int width, int cs, int ucs, int ol, int batch);
int nind_host_sync(nind_net* net);

/* One rank's share of an image on the same pipeline (multi-GPU host entry): crops [crop_begin, crop_end)
 * only.  The image rows those crops read are uploaded step by step, the band they produce is stitched
 * into a device image of full [3][H][W] layout (*d_out, valid until the second next call on this net), and
 * band rows [d2h_y0, d2h_y1) are copied to out_chw_host (full layout too) as they complete.  No
 * synchronisation: nind_host_join() orders another stream (NCCL seam exchange, further copies) after the
 * pipeline's compute stream, nind_host_sync() waits for everything. */
int nind_tiled_denoise_host_range(nind_net* net, const float* img_chw_host, float* out_chw_host, int height,
                                  int width, int cs, int ucs, int ol, int batch, int crop_begin, int crop_end,
                                  int d2h_y0, int d2h_y1, float** d_out);
int nind_host_join(nind_net* net, void* stream);
/* Same, but only after the step of the LAST nind_tiled_denoise_host_range call that makes band rows [band start, y)
 * final: lets a rank hand the rows a neighbour owns to that neighbour while its remaining crops still run. */
int nind_host_join_rows(nind_net* net, int y, void* stream);
/* rows[k] = first band row that is NOT final after step k of the last nind_tiled_denoise_host_range call. */
int nind_host_rows_done(nind_net* net, int* rows, int max_rows, int* n);

/* Page-lock an existing host range (e.g. a shared-memory mapping every rank of a multi-GPU job writes
 * its output rows into) so that copies to/from it are true asynchronous DMA.  The reference has no
 * counterpart: it moves each crop with a synchronous .cpu() (denoise_image.py:254). */
int nind_host_register(void* ptr, size_t bytes);
int nind_host_unregister(void* ptr);

/* File formats either side of the path, on the current device (src and dst are device pointers):
 *   nind_image_to_chw_f32  replaces the conversion of img_path_to_np_flt (common/libs/np_imgops.py:19-28):
 *                          a decoded interleaved [h,w,3] image (BGR order if bgr != 0, as cv2 returns it) of
 *                          type NIND_PIX_* -> planar RGB fp32 [3,h,w]; u8/255, u16/65535, float unchanged
 *                          (bit-identical to the numpy expressions);
 *   nind_chw_f32_to_image  replaces the quantisation of tensor_to_imgfile (common/libs/pt_helpers.py:24-32):
 *                          NIND_PIX_U16 clip(0,1)*65535 rounded half to even, NIND_PIX_U8 clip(0,1)*255+0.5
 *                          truncated (torchvision.utils.save_image), NIND_PIX_F32 unclamped. */
int nind_image_to_chw_f32(const void* src_hwc, int dtype, int height, int width, int bgr, float* dst_chw,
                          void* stream);
int nind_chw_f32_to_image(const float* src_chw, int height, int width, int dtype, int bgr, void* dst_hwc,
                          void* stream);

/* Number of CUDA kernels this library has launched on the calling process so far. */
int64_t nind_kernel_launches(void);

/* Per-layer device timing of the most recent nind_net_forward with timing enabled.
 * nind_set_timing(net, 1) makes forward record one CUDA event pair per layer (adds sync). */
int nind_set_timing(nind_net* net, int enabled);
int nind_get_layer_times(nind_net* net, int max_layers, const char** names, float* ms, double* flops,
                         int* n_layers);
/* Algorithmic HBM bytes (inputs read once + outputs written once) of the same layers. */
int nind_get_layer_bytes(nind_net* net, int max_layers, double* bytes, int* n_layers);

/* Tuning knobs (affect plans built afterwards): "n_tile_deep" (128|256), "max_ctas",
 * "cta_group" (0 auto | 1 | 2), "fuse_pool" (0|1), "pair64" (0|1: pixel-pair mode of the C_out = 64 3x3
 * layers), "flat" (-1 auto | 0 off: 1-D tiles on narrow maps); host pipeline: "host_first" / "host_last" =
 * crops in its first / last step (-1: up to / from the nearest grid-row boundary). */
int nind_set_option(nind_net* net, const char* key, int value);

const char* nind_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NIND_B200_H */
