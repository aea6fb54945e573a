// PathTracer.cs — drop-in for OpenTK-PathTracer/src/Render/PathTracer.cs: the same public surface (NumSpheres, NumCuboids,
// RayDepth, SPP, FocalLength, ApertureDiameter, EnvironmentMap, Result, Samples, Render, SetSize, ResetRenderer), with the GLSL
// dispatch replaced by libptb200.so (Ptb.* in PtbNative.cs).  `Result` stays the Rgba32f GL texture ScreenEffect.Render samples
// (MainWindow.cs:51): it is registered with the library once through CUDA-GL interop and filled on the device after every
// Render() — the image never crosses PCIe.
//
// Host edits besides swapping this file in (INTEGRATION.md §2):
//   MainWindow.cs:131-132,279   BasicDataUBO.SubData(off, size, x)        ->  PathTracer.BasicDataSubData(off, size, x)
//   BaseSTD140Compatible.cs:15  bufferObject.SubData(BufferOffset, ...)   ->  MainWindow.PathTracer.GameObjectsSubData(BufferOffset, data)
//   MainWindow.cs:174-175,189   AtmosphericScatterer.Render() + env arg   ->  PathTracer.GenerateAtmosphere(...) (or keep the GL pass
//                               and hand its faces over with SetEnvironmentFaces)
// Not compiled in this repository: the build image has no .NET SDK.  tests/test_host.py checks the P/Invoke surface against
// include/ptb200.h; the GPU tests drive the same calls in the same order through ctypes.
using System;
using OpenTK;
using OpenTK.Graphics.OpenGL4;
using OpenTK_PathTracer.Render.Objects;

namespace OpenTK_PathTracer
{
    class PathTracer : IDisposable
    {
        readonly IntPtr ctx;
        int numSpheres, numCuboids, rayDepth, spp;
        float focalLength, apertureDiameter;

        public int NumSpheres { get => numSpheres; set { Ptb.Check(Ptb.ptb_set_num_spheres(ctx, value)); numSpheres = value; } }
        public int NumCuboids { get => numCuboids; set { Ptb.Check(Ptb.ptb_set_num_cuboids(ctx, value)); numCuboids = value; } }
        public int RayDepth { get => rayDepth; set { Ptb.Check(Ptb.ptb_set_ray_depth(ctx, value)); rayDepth = value; } }
        public int SPP { get => spp; set { Ptb.Check(Ptb.ptb_set_spp(ctx, value)); spp = value; } }
        public float FocalLength { get => focalLength; set { Ptb.Check(Ptb.ptb_set_focal_length(ctx, value)); focalLength = value; } }
        public float ApertureDiameter { get => apertureDiameter; set { Ptb.Check(Ptb.ptb_set_aperture_diameter(ctx, value)); apertureDiameter = value; } }

        /// <summary>Kept for source compatibility (Gui.cs:84-86 assigns it).  The library holds its own copy of the environment:
        /// assign through SetEnvironmentFaces / SetSkyBox / GenerateAtmosphere.</summary>
        public Texture EnvironmentMap;

        /// <summary>The accumulation image as the GL texture the post-process pass samples.</summary>
        public readonly Texture Result;

        public PathTracer(Texture environmentMap, int width, int height, int rayDepth, int spp, float focalLength, float apertureDiamater, int cudaDevice = 0)
        {
            Ptb.Check(Ptb.ptb_create(out ctx, width, height, MainWindow.MAX_GAMEOBJECTS_SPHERES, MainWindow.MAX_GAMEOBJECTS_CUBOIDS, cudaDevice));
            Result = new Texture(TextureTarget2d.Texture2D);
            Result.SetFilter(TextureMinFilter.Linear, TextureMagFilter.Linear);
            AllocateResult(width, height);
            RayDepth = rayDepth;
            SPP = spp;
            FocalLength = focalLength;
            ApertureDiameter = apertureDiamater;
            EnvironmentMap = environmentMap;
        }

        void AllocateResult(int width, int height)
        {
            Result.MutableAllocate(width, height, 1, PixelInternalFormat.Rgba32f);
            // the GL context of the window thread is current here (MainWindow.OnLoad / OnResize)
            Ptb.Check(Ptb.ptb_register_gl_texture(ctx, (uint)Result.ID));
        }

        public int Samples => Ptb.ptb_samples(ctx);

        /// <summary>One frame: thisRenderNumFrame++ inside the library, then the image goes into Result on the device.</summary>
        public void Render()
        {
            Ptb.Check(Ptb.ptb_render(ctx));
            Ptb.Check(Ptb.ptb_present_gl(ctx));
            Ptb.Check(Ptb.ptb_synchronize(ctx));      // GL samples Result right after this call (MainWindow.cs:51)
        }

        public void SetSize(int width, int height)
        {
            Ptb.Check(Ptb.ptb_set_size(ctx, width, height));      // frame counter = 0; drops the old registration
            AllocateResult(width, height);
        }

        public void ResetRenderer() => Ptb.Check(Ptb.ptb_reset(ctx));

        // ---- what BufferObject.SubData did for UBO bindings 0 and 1
        public unsafe void BasicDataSubData<T>(int offset, int size, T data) where T : unmanaged
        {
            // MainWindow.cs:132 passes a 12-byte Vector3 with size 16: copy into a 16-byte scratch so the library never reads past it
            byte* scratch = stackalloc byte[64];
            int have = Math.Min(sizeof(T), 64);
            Buffer.MemoryCopy(&data, scratch, 64, have);
            for (int i = have; i < Math.Min(size, 64); i++) scratch[i] = 0;
            Ptb.Check(Ptb.ptb_basic_data_subdata(ctx, offset, Math.Min(size, 64), scratch));
        }

        public unsafe void GameObjectsSubData(int offset, Vector4[] data)
        {
            fixed (Vector4* p = data)
                Ptb.Check(Ptb.ptb_game_objects_subdata(ctx, offset, data.Length * sizeof(Vector4), p));
        }

        // ---- environment
        public unsafe void GenerateAtmosphere(int size, byte[] atmosphericDataUbo, Vector3 lightPos, float lightIntensity, int iSteps, int jSteps)
        {
            float* lp = stackalloc float[3] { lightPos.X, lightPos.Y, lightPos.Z };
            fixed (byte* ubo = atmosphericDataUbo)
                Ptb.Check(Ptb.ptb_generate_atmosphere(ctx, size, ubo, atmosphericDataUbo.Length, lp, lightIntensity, iSteps, jSteps));
        }

        public unsafe void SetEnvironmentFaces(int faceSize, float[] sixFacesRgba32f)
        {
            fixed (float* p = sixFacesRgba32f) Ptb.Check(Ptb.ptb_set_environment_rgba32f(ctx, faceSize, p));
        }

        public unsafe void SetSkyBox(int faceSize, byte[] sixFacesSrgb8)
        {
            fixed (byte* p = sixFacesSrgb8) Ptb.Check(Ptb.ptb_set_environment_srgb8(ctx, faceSize, p));
        }

        /// <summary>PrecisionExact (default, bit-identical to the CPU oracle) or PrecisionFast (MUFU + FMA build).</summary>
        public void SetPrecision(int precision) => Ptb.Check(Ptb.ptb_set_precision(ctx, precision));

        public void Dispose()
        {
            Ptb.ptb_unregister_gl_texture(ctx);
            Ptb.ptb_destroy(ctx);
        }
    }
}
