// Family D (UNet_downscale, DDPM_clean_application/src/unet_ms.py) — kernels and program.  Included by b200ddpm.cu after
// Handle/Builder are defined.
#pragma once
namespace b2d {
static int pack_family_d(Handle* h) { (void)h; return fail(-4, "Family D (UNet_downscale) is not built yet in this round"); }
static int build_program_d(Handle* h, int B) { (void)h; (void)B; return fail(-4, "Family D (UNet_downscale) is not built yet in this round"); }
static int set_conditioning_d(Handle* h, const float* cond, int ch, int cw, int B, cudaStream_t st) {
    (void)h; (void)cond; (void)ch; (void)cw; (void)B; (void)st;
    return fail(-4, "Family D (UNet_downscale) is not built yet in this round");
}
}  // namespace b2d
