// PtbNative.cs — P/Invoke declarations for libptb200.so (include/ptb200.h), one per C entry point the host needs.
// Part of the drop-in for OpenTK-PathTracer/src/Render/PathTracer.cs; see PathTracer.cs in this directory and INTEGRATION.md.
// Not compiled in this repository (no .NET toolchain in the build image); tests/test_host.py checks every name and argument
// count below against the header, and the same call sequence is exercised through ctypes by the GPU tests.
using System;
using System.Runtime.InteropServices;

namespace OpenTK_PathTracer
{
    static class Ptb
    {
        const string Lib = "ptb200";          // libptb200.so next to the executable or on LD_LIBRARY_PATH

        public const int FormatRgba32f = 0, FormatRgb32f = 1, FormatRgba8 = 2;
        public const int PrecisionExact = 0, PrecisionFast = 1;

        [DllImport(Lib)] public static extern IntPtr ptb_last_error();
        [DllImport(Lib)] public static extern int ptb_version();
        [DllImport(Lib)] public static extern int ptb_create(out IntPtr ctx, int width, int height, int maxSpheres, int maxCuboids, int device);
        [DllImport(Lib)] public static extern void ptb_destroy(IntPtr ctx);
        [DllImport(Lib)] public static extern int ptb_set_size(IntPtr ctx, int width, int height);
        [DllImport(Lib)] public static extern int ptb_reset(IntPtr ctx);
        [DllImport(Lib)] public static extern int ptb_set_ray_depth(IntPtr ctx, int rayDepth);
        [DllImport(Lib)] public static extern int ptb_set_spp(IntPtr ctx, int spp);
        [DllImport(Lib)] public static extern int ptb_set_focal_length(IntPtr ctx, float focalLength);
        [DllImport(Lib)] public static extern int ptb_set_aperture_diameter(IntPtr ctx, float apertureDiameter);
        [DllImport(Lib)] public static extern int ptb_set_num_spheres(IntPtr ctx, int n);
        [DllImport(Lib)] public static extern int ptb_set_num_cuboids(IntPtr ctx, int n);
        [DllImport(Lib)] public static extern unsafe int ptb_basic_data_subdata(IntPtr ctx, int offset, int size, void* data);
        [DllImport(Lib)] public static extern unsafe int ptb_game_objects_subdata(IntPtr ctx, int offset, int size, void* data);
        [DllImport(Lib)] public static extern unsafe int ptb_set_environment_rgba32f(IntPtr ctx, int faceSize, float* sixFaces);
        [DllImport(Lib)] public static extern unsafe int ptb_set_environment_srgb8(IntPtr ctx, int faceSize, byte* sixFaces);
        [DllImport(Lib)] public static extern unsafe int ptb_generate_atmosphere(IntPtr ctx, int faceSize, void* ubo, int uboSize, float* lightPos, float lightIntensity, int iSteps, int jSteps);
        [DllImport(Lib)] public static extern int ptb_render(IntPtr ctx);
        [DllImport(Lib)] public static extern int ptb_render_frames(IntPtr ctx, int n);
        [DllImport(Lib)] public static extern int ptb_samples(IntPtr ctx);
        [DllImport(Lib)] public static extern int ptb_set_precision(IntPtr ctx, int precision);
        [DllImport(Lib)] public static extern int ptb_register_gl_texture(IntPtr ctx, uint glTexture);
        [DllImport(Lib)] public static extern int ptb_present_gl(IntPtr ctx);
        [DllImport(Lib)] public static extern int ptb_unregister_gl_texture(IntPtr ctx);
        [DllImport(Lib)] public static extern unsafe int ptb_read_result(IntPtr ctx, float* rgba32f);
        [DllImport(Lib)] public static extern unsafe int ptb_tonemap_rgba8(IntPtr ctx, byte* rgba8);
        [DllImport(Lib)] public static extern int ptb_synchronize(IntPtr ctx);

        public static void Check(int rc)
        {
            if (rc < 0)
                throw new InvalidOperationException($"ptb200 error {rc}: {Marshal.PtrToStringAnsi(ptb_last_error())}");
        }
    }
}
