{-# LANGUAGE ForeignFunctionInterface #-}
-- | Raw bindings to libgudni_b200.so (include/gudni_b200.h).  SOURCE ONLY: GHC is not available
-- in the build image, so this module has not been compiled; it shows the binding a Gudni
-- maintainer adds next to the existing direct FFI in Graphics.Gudni.Interface.GLInterop
-- (src/Graphics/Gudni/Interface/GLInterop.hs:60-113).
module Graphics.Gudni.CUDA.FFI where

import Data.Int (Int64)
import Foreign.C.Types
import Foreign.C.String (CString)
import Foreign.Ptr
import Data.Word

-- | Opaque gudni_ctx.
data GudniCtx

-- | gudni_spec = RasterSpec (OpenCL/Rasterizer.hs:37-50), six CInts in declaration order.
-- | gudni_stats: see include/gudni_b200.h; Storable instances live in Graphics.Gudni.CUDA.Setup.

foreign import ccall safe "gudni_b200_init"
  c_init :: CInt -> Ptr CInt {- want spec or nullPtr -} -> Ptr CInt {- got spec -} -> Ptr (Ptr GudniCtx) -> IO CInt

foreign import ccall safe "gudni_b200_frame_begin"
  c_frameBegin :: Ptr GudniCtx
               -> Ptr CChar -> CSize            -- geoGeometryPile bytes (Raster/Serialize.hs:111)
               -> Ptr CFloat -> CInt            -- suSubstancePile (Raster/Serialize.hs:195)
               -> Ptr Word8 -> CSize            -- picture heap (Figure/Picture.hs:159)
               -> Ptr () -> CInt                -- picture usages, 24 bytes each
               -> Ptr CFloat                    -- suBackgroundColor r g b a
               -> CInt -> CInt -> CInt          -- bitmap width, height, frame count
               -> IO CInt

foreign import ccall safe "gudni_b200_raster_job"
  c_rasterJob :: Ptr GudniCtx
              -> Ptr () -> CInt                 -- rJShapePile (16 bytes each)
              -> Ptr () -> CInt                 -- rJTilePile  (32 bytes each)
              -> CInt -> CInt                   -- rJColumnAllocation, jobIndex
              -> IO CInt

foreign import ccall safe "gudni_b200_raster_scene"
  c_rasterScene :: Ptr GudniCtx -> Ptr () -> CInt -> IO CInt   -- un-binned shape entries (32 bytes each)

-- | Level 3: the scene before serialisation.  Replaces onShape's geometry work (Raster/Serialize.hs:148-177),
-- enclose (Raster/Enclosure.hs:62-73) and outlineToStrands (Raster/Strand.hs:153-178); call c_frameBegin
-- with a null geometry pile.  Records: shape 32 bytes (tag, outline slice, transform slice), outline 8 bytes
-- (pair slice), curve pair 16 bytes (onCurve, offCurve), transform 16 bytes (kind, a, b) — see
-- include/gudni_b200.h.
foreign import ccall safe "gudni_b200_raster_outlines"
  c_rasterOutlines :: Ptr GudniCtx
                   -> Ptr () -> CInt            -- shapes as traverseShapeTree visits them (first = top-most)
                   -> Ptr () -> CInt            -- outlines (shared between shapes: one glyph, many placements)
                   -> Ptr CFloat -> Int64       -- curve pairs
                   -> Ptr () -> CInt            -- simple transformations, outermost first per shape
                   -> IO CInt

foreign import ccall safe "gudni_b200_frame_end"
  c_frameEnd :: Ptr GudniCtx -> Ptr CUInt {- HostBitmapTarget pointer, DrawTarget.hs:34 -} -> Ptr () -> IO CInt

foreign import ccall safe "gudni_b200_last_error"
  c_lastError :: Ptr GudniCtx -> IO CString

foreign import ccall safe "gudni_b200_destroy"
  c_destroy :: Ptr GudniCtx -> IO ()

-- | Optional: page-lock a long-lived caller buffer (a Pile's allocation, the HostBitmapTarget) so the
-- per-frame copies run at full PCIe rate; unregister before freeing / after a Pile has grown.
foreign import ccall safe "gudni_b200_host_register"
  c_hostRegister :: Ptr GudniCtx -> Ptr () -> CSize -> IO CInt

foreign import ccall safe "gudni_b200_host_unregister"
  c_hostUnregister :: Ptr GudniCtx -> Ptr () -> IO CInt

-- | Optional: make a page-locked HostBitmapTarget the frame's target, so that the kernels store their rows into it
-- across PCIe while the rest of the frame is rasterized and c_frameEnd (given the same pointer) copies nothing.
-- nullPtr restores the library's own frame buffer.
foreign import ccall safe "gudni_b200_frame_target_host"
  c_frameTargetHost :: Ptr GudniCtx -> Ptr CUInt -> IO CInt

-- | Optional (multi-GPU hosts, one process per device): restrict the frame to whole root-tile rows
-- [rowBegin, rowEnd) of the canvas; call between frame_begin and the raster calls.
foreign import ccall safe "gudni_b200_frame_strip"
  c_frameStrip :: Ptr GudniCtx -> CInt -> CInt -> IO CInt

-- | gudni_generations (five Word64: geometry, substances, pictures, picture uses, entries).  A generation that is
-- unchanged since the last frame means "the bytes behind this pointer are the ones already on the device": the upload is
-- skipped.  0 = always upload.  The caller bumps a counter whenever it refills the Pile (resetPile / addToPile).
foreign import ccall safe "gudni_b200_frame_begin_cached"
  c_frameBeginCached :: Ptr GudniCtx
                     -> Ptr CChar -> CSize -> Ptr CFloat -> CInt -> Ptr Word8 -> CSize -> Ptr () -> CInt
                     -> Ptr CFloat -> CInt -> CInt -> CInt
                     -> Ptr Word64               -- gudni_generations
                     -> IO CInt

foreign import ccall safe "gudni_b200_raster_scene_cached"
  c_rasterSceneCached :: Ptr GudniCtx -> Ptr () -> CInt -> Word64 -> IO CInt

-- | Several GPUs of one box behind one call (include/gudni_b200.h, gudni_b200_multi_*): the canvas is cut into strips
-- of whole root-tile rows, one per device, re-cut every frame from the measured times; every device copies its rows
-- straight into the HostBitmapTarget.
data GudniMulti

foreign import ccall safe "gudni_b200_multi_init"
  c_multiInit :: CInt -> Ptr CInt {- devices or nullPtr -} -> Ptr CInt {- want spec or nullPtr -} -> Ptr CInt {- got spec -}
              -> Ptr (Ptr GudniMulti) -> IO CInt

foreign import ccall safe "gudni_b200_multi_frame"
  c_multiFrame :: Ptr GudniMulti
               -> Ptr CChar -> CSize -> Ptr CFloat -> CInt -> Ptr Word8 -> CSize -> Ptr () -> CInt
               -> Ptr CFloat -> CInt -> CInt -> CInt
               -> Ptr () -> CInt                 -- un-binned shape entries (32 bytes each)
               -> Ptr Word64                     -- gudni_generations or nullPtr
               -> Ptr CUInt                      -- HostBitmapTarget pointer
               -> Ptr ()                         -- gudni_multi_stats or nullPtr
               -> IO CInt

foreign import ccall safe "gudni_b200_multi_last_error"
  c_multiLastError :: Ptr GudniMulti -> IO CString

foreign import ccall safe "gudni_b200_multi_destroy"
  c_multiDestroy :: Ptr GudniMulti -> IO ()
