/*
 * msi_b200.h -- C ABI of the Blackwell-native multi-sphere-image (MSI) inference path.
 *
 * The reference (brownvc/matryodshka @ 831c407) is pure Python/TensorFlow and has
 * NO native boundary of its own: the path sits behind the
 * Python call surface matryodshka.msi.MSI / geometry.projector.  This header is
 * the boundary a binding for that surface needs: each entry point replaces the
 * graph of stock TF ops behind one reference function, cited as file:line of
 * /root/reference.  INTEGRATION.md shows the ctypes stub the Python side uses.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (cudaMalloc'd / torch CUDA storage) unless
 *     the parameter name ends in _host; tensors are dense NHWC float32 as in the
 *     reference (images [B,H,W,3], PSV [B,H,W,6P], RGBA layers [B,H,W,L,4]);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls
 *     only enqueue work, they never synchronise; reentrant per stream;
 *   - return value: 0 = MSI_OK, <0 = error (msi_last_error() gives the text);
 *     asynchronous CUDA faults surface at the caller's next synchronisation;
 *   - no entry point falls back to the CPU: without a CUDA device every compute
 *     call returns MSI_ERR_CUDA.
 */
#ifndef MSI_B200_H_
#define MSI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSI_B200_ABI_VERSION 1

#define MSI_OK 0
#define MSI_ERR_INVALID_ARG (-1)
#define MSI_ERR_CUDA (-2)
#define MSI_ERR_UNSUPPORTED (-3)
#define MSI_ERR_STATE (-4)

/* image element types accepted by msi_psv_build */
#define MSI_IMG_F32 0 /* float32 in [0,1] (or already in [-1,1] when preprocess=0) */
#define MSI_IMG_U8 1  /* uint8 in [0,255]: x/255 as tf.image.convert_image_dtype  */

/* convolution back ends of the net (both are hand-written kernels of this library) */
#define MSI_CONV_TCGEN05 0 /* tcgen05.mma + TMA implicit GEMM, accumulators in TMEM   */
#define MSI_CONV_SIMT 1    /* fp32 CUDA-core implicit GEMM: bring-up / cross-check    */

/* operand precision of the tcgen05 back end */
#define MSI_PREC_FP16X3 0 /* fp16 hi/lo split, 3 MMAs per product: ~fp32 accuracy (default; meets 1e-3) */
#define MSI_PREC_FP16 1   /* single fp16 MMA: 3x fewer MMAs, ~7e-3 max-abs on the net output  */
#define MSI_PREC_FP16_FP8X 2 /* fp16 main product + both cross terms as ONE e4m3 MMA of twice the K, on the layers
                              * whose N tile is 128 (every conv / deconv with Cout >= 128); the Cout = 64 layers and the
                              * head keep MSI_PREC_FP16X3.  2 MMA units per product instead of 3; ~2.5e-4 max-abs on
                              * the net output (meets 1e-3) */

int msi_b200_abi_version(void);
const char* msi_last_error(void);

/* Number of CUDA kernels this library has launched in this process (all entry
 * points, all threads).  bench.py reports the delta over its timed region. */
uint64_t msi_launch_count(void);

/* ------------------------------------------------------------------------- *
 * Stage 1 -- spherical plane-sweep volume
 * replaces: MSI.format_network_input (matryodshka/msi.py:1094-1130) ->
 *   projector.ods_sphere_sweep / sweep_one (geometry/projector.py:129-170,209-211) ->
 *   spherical.lat_long_grid (:42-44), backproject_spherical (:116-129),
 *   projector.apply_pose (:275-291), spherical.project_ods (:170-233),
 *   sampling.bilinear_wrapper2 / resample (geometry/sampling.py:59-67,135-197),
 *   and MSI.preprocess_image (msi.py:1163-1171) when preprocess=1.
 *
 * ref, src      [B,H,W,3] images (eye 0 = ref, order +1; eye 1 = src, order -1)
 * poses         [B,2,16]  row-major 4x4 `pose_eye . ref_pose_inv` per frame and eye
 * baselines     [B]       ODS baseline radius = intrinsics[b,0,0] (data_loader.py:160)
 * depths        [P]       sphere radii, far -> near (MSI.inv_depths, msi.py:1196)
 * cos_s,sin_s   [W], cos_t,sin_t [H]: cos/sin of the pixel-centre longitudes /
 *               latitudes of lat_long_grid (the "ERP-coord tables")
 * out_f32       [B,H,W,6P] float32 PSV, channel = eye*3P + p*3 + rgb   (may be NULL)
 * out_hi,out_lo [B,H,W,c_stride] fp16 hi/lo split of the PSV scaled by
 *               MSI_ACT_SCALE, the operand format of the conv net (may be NULL);
 *               c_stride >= 6P, channels [6P, c_stride) are zero-filled.
 * ------------------------------------------------------------------------- */
#define MSI_ACT_SCALE 16.0f
#define MSI_WEIGHT_SCALE 1024.0f

int msi_psv_build(const void* ref, const void* src, int img_dtype, int preprocess,
                  const float* poses, const float* baselines, const float* depths,
                  const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t,
                  int B, int H, int W, int P,
                  float* out_f32, void* out_hi, void* out_lo, int c_stride,
                  void* scratch, size_t scratch_bytes, void* stream);

/* Optional scratch for msi_psv_build (16-byte aligned device memory, >= this many
 * bytes): with it the sweep first packs both images to preprocessed RGBX float4
 * and evaluates both eyes per thread (2 launches, ~2x faster); with scratch ==
 * NULL it runs the single-kernel form.  Results are identical. */
size_t msi_psv_scratch_bytes(int B, int H, int W);

/* ------------------------------------------------------------------------- *
 * Stage 1 with cached sweep coordinates (static rig).
 * The sample coordinates of the sweep depend on (poses, baselines, depths, H, W)
 * only -- not on the images -- and the reference's data has one rig for a whole
 * sequence (identity eye poses, one baseline: data_loader.py:146-174), so the
 * chain backproject_spherical -> apply_pose -> project_ods (spherical.py:116-129,
 * projector.py:275-291, spherical.py:170-233) is evaluated ONCE into a table and
 * every frame is a pure gather + bilinear blend (sampling.py:135-197) from it.
 * The table is written by the same device functions as msi_psv_build, so
 * msi_psv_gather returns the same bits as msi_psv_build (validity mask included:
 * an invalid sample carries its (1,1)).  Rebuild the table when a pose, a
 * baseline or the depths change; the jittered sweep (msi.py:1118-1120) keeps
 * msi_psv_build.
 *
 * table   [frames][H][W][P] float4 = (u_ref, v_ref, u_src, v_src), 16-byte aligned
 *         device memory of msi_sweep_table_bytes(frames, H, W, P) bytes; frames = 1
 *         (one rig shared by every frame of a batch) or B (poses [frames,2,16],
 *         baselines [frames]).
 * msi_psv_gather: table_frames must be 1 or B; other arguments as msi_psv_build.
 * ------------------------------------------------------------------------- */
size_t msi_sweep_table_bytes(int frames, int H, int W, int P);
int msi_sweep_table_build(const float* poses, const float* baselines, const float* depths,
                          const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t,
                          int frames, int H, int W, int P, void* table, void* stream);
int msi_psv_gather(const void* ref, const void* src, int img_dtype, int preprocess,
                   const void* table, int table_frames, int B, int H, int W, int P,
                   float* out_f32, void* out_hi, void* out_lo, int c_stride,
                   void* scratch, size_t scratch_bytes, void* stream);

/* Sample coordinates only (spherical.project_ods, spherical.py:170-233, after
 * backproject + apply_pose): uv [B,2,P,H,W,2] (x=u, y=v) and valid [B,2,P,H,W]
 * (uint8, 0 where disc < 0 and the sample snaps to pixel (1,1), :226-229). */
int msi_sweep_coords(const float* poses, const float* baselines, const float* depths,
                     const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t,
                     int B, int H, int W, int P, float* uv, uint8_t* valid, void* stream);

/* ------------------------------------------------------------------------- *
 * Stage 2 1/2 -- RGBA layer assembly, `blend_psv`
 * replaces: MSI.infer_msi layer_prediction block (msi.py:130-147).
 * pred [B,H,W,2L] net output in (-1,1); PSV either float32 (psv_f32) or the
 * fp16 hi/lo pair (psv_f32 == NULL), channel stride c_stride.
 * rgba [B,H,W,L,4]; blend_weights / alphas [B,H,W,L] optional (NULL to skip).
 * ------------------------------------------------------------------------- */
int msi_rgba_assemble(const float* pred, const float* psv_f32, const void* psv_hi, const void* psv_lo,
                      int c_stride, int B, int H, int W, int L,
                      float* rgba, float* blend_weights, float* alphas, void* stream);

/* The other colour-prediction schemes of infer_msi (matryodshka/msi.py:107-116, 166-273).  pred has
 * n_pred channels per pixel:
 *   MSI_COLOR_BLEND_PSV     2L     [w | alpha]                      msi.py:117-147 (= msi_rgba_assemble)
 *   MSI_COLOR_BLEND_BG      2L+3   [w | alpha | bg rgb]             msi.py:166-190
 *   MSI_COLOR_BLEND_BG_PSV  3L+3   [w | alpha | bg_w | bg rgb]      msi.py:209-247
 *   MSI_COLOR_ALPHA_ONLY    L      [alpha], rgb = reference-eye PSV  msi.py:249-268
 * blend_weights / alphas / bg_blend_weights [B,H,W,L] are optional outputs (msi.py:276-286). */
#define MSI_COLOR_BLEND_PSV 0
#define MSI_COLOR_BLEND_BG 1
#define MSI_COLOR_BLEND_BG_PSV 2
#define MSI_COLOR_ALPHA_ONLY 3
int msi_rgba_assemble_ex(const float* pred, int n_pred, const float* psv_f32, const void* psv_hi, const void* psv_lo,
                         int c_stride, int B, int H, int W, int L, int mode, float* rgba, float* blend_weights,
                         float* alphas, float* bg_blend_weights, void* stream);
/* Same, with the prediction's pixel stride given separately (pred_stride >= n_pred floats): reads the
 * tensor-core net's output in place when its head is padded to a multiple of 64 channels. */
int msi_rgba_assemble_strided(const float* pred, int n_pred, int pred_stride, const float* psv_f32, const void* psv_hi,
                              const void* psv_lo, int c_stride, int B, int H, int W, int L, int mode, float* rgba,
                              float* blend_weights, float* alphas, float* bg_blend_weights, void* stream);

/* ------------------------------------------------------------------------- *
 * Stage 3 -- reproject the L spheres to the target position and over-composite
 * replaces: MSI.msi_render_equirect_view (msi.py:407-429) and
 *   msi_render_equirect_depth (:384-405) in ONE pass ->
 *   projector.projective_forward_sphere (projector.py:34-62),
 *   spherical.intersect_sphere (spherical.py:268-326), project_spherical
 *   (:235-246), theta_phi_to_pixels (:54-68), sampling.resample,
 *   projector.over_composite (:246-265), over_composite_depth (:225-244),
 *   MSI.deprocess_image / deprocess_depth_image (msi.py:1173-1194).
 *
 * rgba        [B,H,W,L,4]
 * tgt_pose_rt [B,16] row-major [R|t]; tgt_pos [B,3] target offset (x,y,z)
 * depths      [L]
 * out_rgb     [B,H,W,3] float32 in [-1,1]       (NULL to skip)
 * out_depth   [B,H,W,3] float32 in [0,1)        (NULL to skip)
 * out_rgb_u8 / out_depth_u8 [B,H,W,3] uint8, convert_image_dtype semantics
 *
 * Arithmetic contract.  The render side has no validity mask, so the fused kernel
 * evaluates the per-sample chain (quadratic root, two atan2, pixel scaling) with
 * approximate reciprocal / square root, FMA and a polynomial atan2 (2.5e-7 rad):
 * sample coordinates agree with the reference's float32 chain to ~1e-4 px
 * (contract: 1e-3 px; floor() differs only at knife-edge coordinates), the
 * bilinear blend and the over-composite recurrence keep the reference's operation
 * order without FMA.  msi_intersect_sphere_coords_ex(fast=1) returns exactly the
 * coordinates this kernel samples at; msi_project_layers / msi_intersect_sphere_coords
 * keep the strict chain.  Environment MSI_RENDER_V1=1 selects the strict-chain form
 * of the fused kernel (A/B runs).
 * ------------------------------------------------------------------------- */
int msi_render_composite(const float* rgba, const float* tgt_pose_rt, const float* tgt_pos,
                         const float* depths,
                         const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t,
                         int B, int H, int W, int L,
                         float* out_rgb, float* out_depth, uint8_t* out_rgb_u8, uint8_t* out_depth_u8,
                         void* stream);

/* msi_render_composite fused with the path's only collective (one all-gather of the
 * rendered frames).  Besides the local outputs, the uint8 view of local frame b is stored into the
 * gathered buffer [world * B, H, W, 3] of EVERY rank at frame index first_frame + b:
 *   peer_rgb_u8       device array of n_peers base pointers of that buffer on each rank (peer-mapped
 *                     symmetric memory; includes this rank), or NULL for no gather;
 *   multicast_rgb_u8  the NVSwitch multicast address of the same buffer, or NULL: when given, one
 *                     multimem.st per 32-bit word replaces the n_peers stores.
 * No collective kernel runs afterwards; readers synchronise across ranks (a barrier) before use. */
int msi_render_composite_gather(const float* rgba, const float* tgt_pose_rt, const float* tgt_pos, const float* depths,
                                const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t, int B,
                                int H, int W, int L, float* out_rgb, float* out_depth, uint8_t* out_rgb_u8,
                                uint8_t* out_depth_u8, uint8_t* const* peer_rgb_u8, int n_peers,
                                uint8_t* multicast_rgb_u8, long long first_frame, void* stream);

/* The MSI seen from one ODS eye: MSI.msi_render_ods_view (msi.py:502-525) ->
 * projector.projective_forward_ods (projector.py:101-127) -> spherical.intersect_ods
 * (spherical.py:328-365) + over_composite.  pose_rt [B,16] (the "jitter pose"), order = +1 (left /
 * ref eye) or -1 (right / src eye), baselines [B]; out_rgb [B,H,W,3] float32 and/or uint8. */
int msi_render_ods(const float* rgba, const float* pose_rt, float order, const float* baselines,
                   const float* depths,
                   const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t,
                   int B, int H, int W, int L, float* out_rgb, uint8_t* out_rgb_u8, void* stream);

/* MSI.msi_render_perspective_view (matryodshka/msi.py:475-500) ->
 * projector.projective_forward_sphere_to_perspective (geometry/projector.py:64-99) ->
 * spherical.intersect_perspective (geometry/spherical.py:367-401) + over_composite: the MSI seen
 * through a pinhole window of oH x oW pixels.  pose_rt [B,16] = the viewing-window rotation the
 * reference builds from viewing_window * pi / 2 about y (projector.py:80-85; the caller's pose is
 * ignored there); tgt_pos [B,3]; s_axis [oW], t_axis [oH] = the uv_grid axes (spherical.py:46-48).
 * out_rgb [B,oH,oW,3] float32 in [-1,1] and / or out_rgb_u8. */
int msi_render_perspective(const float* rgba, const float* pose_rt, const float* tgt_pos, const float* depths,
                           const float* s_axis, const float* t_axis, int B, int H, int W, int L, int oH, int oW,
                           float* out_rgb, uint8_t* out_rgb_u8, void* stream);

/* Sample coordinates only (spherical.intersect_sphere): uv [B,L,H,W,2]. */
int msi_intersect_sphere_coords(const float* tgt_pose_rt, const float* tgt_pos, const float* depths,
                                const float* cos_s, const float* sin_s, const float* cos_t,
                                const float* sin_t, int B, int H, int W, int L, float* uv, void* stream);

/* fast = 0: as above; fast = 1: the coordinates of the fast chain of the fused render kernel. */
int msi_intersect_sphere_coords_ex(const float* tgt_pose_rt, const float* tgt_pos, const float* depths,
                                   const float* cos_s, const float* sin_s, const float* cos_t,
                                   const float* sin_t, int B, int H, int W, int L, int fast, float* uv, void* stream);

/* Reprojected layers without compositing (MSI.msi_render_equirect_view_single,
 * msi.py:431-452): out [L,B,H,W,4]. */
int msi_project_layers(const float* rgba, const float* tgt_pose_rt, const float* tgt_pos,
                       const float* depths,
                       const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t,
                       int B, int H, int W, int L, float* out, void* stream);

/* Point-wise forms of geometry/spherical.py + projector.apply_pose (stage-level API).
 *   MSI_OP_BACKPROJECT_SPHERICAL (spherical.py:116-129): a = S, b = T [n], c = depths [planes]
 *        -> o0, o1, o2 = x, y, z [planes, n]
 *   MSI_OP_APPLY_POSE (projector.py:275-291): a, b, c = x, y, z [planes, n]; pose [16] or [planes,16]
 *        -> o0, o1, o2
 *   MSI_OP_PROJECT_ODS (spherical.py:170-233): a, b, c = x, y, z [n]; order = +1 / -1; baseline
 *        -> o0 = uv [n, 2], valid [n] (optional)
 *   MSI_OP_PROJECT_SPHERICAL (spherical.py:235-246): a, b, c = x, y, z [n] -> o0 = uv [n, 2]
 *   MSI_OP_THETA_PHI_TO_PIXELS (spherical.py:54-68): a = theta, b = phi [n] -> o0 = uv [n, 2]
 * H, W = ERP size used by the pixel mappings. */
#define MSI_OP_BACKPROJECT_SPHERICAL 0
#define MSI_OP_APPLY_POSE 1
#define MSI_OP_PROJECT_ODS 2
#define MSI_OP_PROJECT_SPHERICAL 3
#define MSI_OP_THETA_PHI_TO_PIXELS 4
int msi_point_op(int op, const float* a, const float* b, const float* c, long long n, int planes,
                 const float* pose, int pose_per_plane, float order, float baseline, int H, int W,
                 float* o0, float* o1, float* o2, uint8_t* valid, void* stream);

/* sampling.resample (sampling.py:135-197): image [N,H,W,C], coords [N,h,w,2] -> out [N,h,w,C];
 * bilinear, floor-mod wrap-around in x and y. */
int msi_resample(const float* image, const float* coords, int N, int H, int W, int C, int h, int w,
                 float* out, void* stream);

/* projector.over_composite (:246-265) / over_composite_depth (:225-244):
 * layers [L,B,H,W,4] back to front -> out [B,H,W,3]. */
int msi_over_composite(const float* layers, int L, int B, int H, int W, int depth_mode, float* out,
                       void* stream);

/* ------------------------------------------------------------------------- *
 * High-res plane-streamed re-render (reference driver test.py:284-394), one PSV
 * plane at a time so that a 4096x2048 sweep never holds more than one layer:
 *
 * msi_highres_plane: the high-res PSV plane of both eyes for ONE depth
 *   (format_network_input with a single plane, test.py:312-317), the plane's
 *   low-res blend weight / alpha upsampled bilinearly with align_corners
 *   (:319-325) and the blend (:327-334) -> rgba [Hh,Wh,4].
 *   hres_ref/src [1,Hh,Wh,3]; poses [2,16]; baseline [1]; depth [1];
 *   tables for the Hh x Wh grid; blend_weights / alphas [1,lh,lw,L] (low-res net
 *   outputs), plane = index into L.
 * msi_highres_composite: reproject that layer to the target position
 *   (msi_render_equirect_view_single, :338) and fold it into the running
 *   composite in place (:374-382): acc_rgb, acc_depth [Hh,Wh,3] float32; plane 0
 *   initialises them.
 * ------------------------------------------------------------------------- */
int msi_highres_plane(const void* hres_ref, const void* hres_src, int img_dtype, int preprocess,
                      const float* poses, const float* baseline, const float* depth,
                      const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t,
                      int Hh, int Wh, const float* blend_weights, const float* alphas,
                      int lh, int lw, int L, int plane, float* rgba, void* stream);
int msi_highres_composite(const float* rgba, const float* tgt_pose_rt, const float* tgt_pos,
                          const float* depth,
                          const float* cos_s, const float* sin_s, const float* cos_t, const float* sin_t,
                          int Hh, int Wh, int plane, int num_planes, float* acc_rgb, float* acc_depth,
                          void* stream);

/* ------------------------------------------------------------------------- *
 * Stage 2 -- the conv net
 * replaces: nets.msi_coord_train_net (matryodshka/nets.py:471-515): 14 coord
 * convs 3x3 + LayerNorm + ReLU, 3 transposed convs 4x4 s2 + LayerNorm + ReLU,
 * 1x1 head + bias + tanh (slim.conv2d / conv2d_transpose / layer_norm).
 *
 * A net object owns nothing but small host-side bookkeeping and TMA descriptors;
 * device memory comes from the caller (workspace for activations, arena for
 * packed weights) so that the host framework's allocator stays in charge.
 * ------------------------------------------------------------------------- */
typedef struct msi_net msi_net;

int msi_net_create(msi_net** net, int H, int W, int c_in, int c_out, int ngf, int max_batch,
                   int conv_impl, int precision);

/* Net variants.
 * MSI_NET_COORD  nets.msi_coord_train_net (matryodshka/nets.py:471-515): SAME zero padding, one
 *                |sin(latitude)| coord channel appended to every 3x3 conv input (nets.py:260-270).
 * MSI_NET_WRAP   nets.msi_train_net (matryodshka/nets.py:387-469): no coord channel; every conv /
 *                deconv input goes through wrap_pad (nets.py:288-295) = circular padding along the
 *                width (longitude) and zero padding along the height, then a VALID conv; the
 *                stride-2 convs therefore pad (1, 1) instead of TF-SAME's (0, 1), and the deconvs
 *                are wrap_pad(x, 2, 2) + VALID 4x4 s2 cropped [5:-5] (= the SAME deconv taps read
 *                from the circularly extended input).  tcgen05 back end only.
 * msi_net_create(...) == msi_net_create_ex(..., MSI_NET_COORD). */
#define MSI_NET_COORD 0
#define MSI_NET_WRAP 1
int msi_net_create_ex(msi_net** net, int H, int W, int c_in, int c_out, int ngf, int max_batch,
                      int conv_impl, int precision, int variant);
void msi_net_destroy(msi_net* net);
size_t msi_net_workspace_bytes(const msi_net* net);
size_t msi_net_arena_bytes(const msi_net* net);
int msi_net_input_c_stride(const msi_net* net);
int msi_net_bind(msi_net* net, void* workspace, size_t workspace_bytes, void* arena, size_t arena_bytes);

/* Pack one layer's parameters from the TF checkpoint layout into
 * the arena: conv weights HWIO [k,k,Cin(+1 coord),Cout], deconv [k,k,Cout,Cin],
 * LayerNorm gamma/beta [Cout], head bias [Cout] (NULL where absent). */
int msi_net_load_layer(msi_net* net, const char* scope, const float* weights, const float* gamma,
                       const float* beta, const float* bias, void* stream);

/* Forward B <= max_batch frames.  Input is the PSV either as float32
 * [B,H,W,c_in] (in_f32) or as the fp16 hi/lo pair written by msi_psv_build
 * (in_f32 == NULL; channel stride msi_net_input_c_stride()).  pred [B,H,W,c_out]. */
int msi_net_forward(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo, int B,
                    float* pred, void* stream);

/* msi_net_forward with the RGBA assembly of `blend_psv` (MSI.infer_msi layer_prediction block,
 * matryodshka/msi.py:130-147) fused into the head's epilogue: the 1x1 conv + bias + tanh of
 * nets.py:509-515 never leaves the SM as `pred`; the epilogue turns a pixel's L blend weights and L
 * alphas into rgba [B,H,W,L,4], reading the pixel's two PSV eyes from the net's own input operand
 * (fp16 hi + lo).  Same arithmetic, in the same order, as msi_net_forward + msi_rgba_assemble on
 * that operand (bit for bit).  Available when msi_net_can_fuse_rgba() returns 1: tcgen05 back
 * end, MSI_PREC_FP16X3, c_in = 6P and c_out = 2L with L = P = 32 or 64. */
int msi_net_can_fuse_rgba(const msi_net* net);
int msi_net_forward_rgba(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo, int B,
                         float* rgba, void* stream);

/* The workspace copy of the network input (fp16 hi/lo, channel stride
 * msi_net_input_c_stride()): let msi_psv_build write the PSV there and pass the
 * same pointers to msi_net_forward to skip the copy. */
int msi_net_input_buffers(msi_net* net, void** hi, void** lo);

/* Test hook: copy the post-LayerNorm+ReLU activation of a layer (fp16 hi+lo
 * recombined, un-scaled) into out [B,h,w,C] float32, after a forward. */
int msi_net_read_activation(msi_net* net, const char* scope, int B, float* out, void* stream);
/* Test hook: raw (pre-LayerNorm) conv output of a layer, [B,h,w,C] float32. */
int msi_net_read_raw(msi_net* net, const char* scope, int B, float* out, void* stream);
int msi_net_num_launches_per_forward(const msi_net* net);

/* Measurement hooks (bench.py): the same forward with CUDA events recorded on
 * `stream` around every conv launch and every LayerNorm group.  Synchronises the
 * stream; conv_ms_host / ln_ms_host are HOST arrays of msi_net_num_layers()
 * floats.  msi_net_layer_flops = algorithmic FLOPs of layer i for one frame
 * (2 x MACs, coord channels counted: SURVEY.md 8a a10 table). */
int msi_net_forward_profiled(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo,
                             int B, float* pred, void* stream, float* conv_ms_host, float* ln_ms_host);
/* Same, with two options: `flush` (device buffer larger than the L2, or NULL) is overwritten before every timed
 * launch and each launch gets its own event pair, so the durations are cold-L2 (inputs come from HBM, as in the frame
 * for every tensor larger than the L2); `rgba` (or NULL) runs the head with the fused RGBA assembly
 * (msi_net_forward_rgba), `pred` is then unused and may be NULL. */
int msi_net_forward_profiled_flush(msi_net* net, const float* in_f32, const void* in_hi, const void* in_lo,
                                   int B, float* pred, void* stream, float* conv_ms_host, float* ln_ms_host,
                                   void* flush, size_t flush_bytes, float* rgba);
int msi_net_num_layers(const msi_net* net);
const char* msi_net_layer_scope(const msi_net* net, int i);
double msi_net_layer_flops(const msi_net* net, int i);

#ifdef __cplusplus
}
#endif
#endif /* MSI_B200_H_ */
