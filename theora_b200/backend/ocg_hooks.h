/* ocg_hooks.h -- the reference-side selector for the B200 back-end.
 *
 * Force-included (gcc -include) when the UNMODIFIED reference host sources are
 * compiled.  It does what lib/x86/x86int.h:22-49, lib/arm/armint.h:27-48 and
 * lib/arm/armenc.h:22-25 do for their ISAs: route the accel macros of
 * lib/state.h:62-112, lib/encint.h:57-132 and lib/decint.h:39-46 through the
 * runtime vtables and name this back-end's init functions.  A maintainer
 * adopting the back-end in-tree would instead add
 *     # if defined(OC_B200_CUDA)
 *     #  include "b200/ocg_hooks.h"
 *     # endif
 * next to the OC_X86_ASM / OC_ARM_ASM / OC_C64X_ASM blocks at state.h:48-60,
 * encint.h:46-55 and decint.h:35-37.
 */
#ifndef OCG_HOOKS_H
#define OCG_HOOKS_H

#define OC_STATE_USE_VTABLE (1)
#define OC_ENC_USE_VTABLE   (1)
#define OC_DEC_USE_VTABLE   (1)

#define oc_state_accel_init oc_state_accel_init_ocg
#define oc_enc_accel_init   oc_enc_accel_init_ocg
#define oc_dec_accel_init   oc_dec_accel_init_ocg

struct oc_theora_state;
struct th_enc_ctx;
struct th_dec_ctx;
void oc_state_accel_init_ocg(struct oc_theora_state *_state);
void oc_enc_accel_init_ocg(struct th_enc_ctx *_enc);
void oc_dec_accel_init_ocg(struct th_dec_ctx *_dec);

#endif
