-- | Drop-in for Graphics.Gudni.OpenCL.CallKernels (src/Graphics/Gudni/OpenCL/CallKernels.hs):
-- same exported names and types, bodies call the C ABI instead of CLUtil.  SOURCE ONLY (no GHC in
-- the build image).  buildRasterJobs is unchanged from the reference (:244-255) and not repeated.
module Graphics.Gudni.CUDA.CallKernels
  ( queueRasterJobs
  , queueRasterScene
  , queueRasterSceneMulti
  , ShapeEntryRecord(..)
  ) where

import Graphics.Gudni.CUDA.FFI
import Graphics.Gudni.CUDA.Setup (Rasterizer(..), checkStatus)
import Graphics.Gudni.Raster.Job
import Graphics.Gudni.Raster.Serialize
import Graphics.Gudni.Figure.Picture (makePictData)
import Graphics.Gudni.Interface.DrawTarget
import Graphics.Gudni.Util.Pile

import Control.Lens
import Control.Monad
import Control.Monad.State
import Foreign.Marshal.Array (withArray)
import Foreign.C.String (peekCString)
import Linear (V2(..))
import Foreign.Ptr
import Foreign.C.Types
import Foreign.Storable
import Data.Word (Word64)
import Graphics.Gudni.Figure (Box, leftSide, topSide, rightSide, bottomSide, SubSpace)
import Graphics.Gudni.Raster.Types (ShapeTag, GeoReference(..), NumStrands)

-- | queueRasterJobs (OpenCL/CallKernels.hs:218-242): frame constants once, then every job, then the
-- read-back into the SDL pointer that OutputPtr did (OpenCL/Instances.hs:60-75).
queueRasterJobs :: (MonadIO m, Show token)
                => CInt -> RasterParams token -> [RasterJob] -> GeometryMonad m ()
queueRasterJobs frameCount params jobs = liftIO $ do
    let ctx      = rasterCtx (params ^. rpDevice)
        geoPile  = params ^. rpGeometryState  . geoGeometryPile
        subPile  = params ^. rpSubstanceState . suSubstancePile
        Color' r g b a = colorComponents (params ^. rpSubstanceState . suBackgroundColor)
        P (Point2 w h) = P (targetArea (params ^. rpTarget))
    (pictData, pictUsage) <- makePictData (params ^. rpSubstanceState . suPictureMapping)
                                          (params ^. rpSubstanceState . suPictureUsages)
    withArray [r, g, b, a] $ \bg ->
      checkStatus ctx =<< c_frameBegin ctx
          (castPtr (geoPile ^. pileData))  (fromIntegral (geoPile ^. pileCursor))
          (castPtr (subPile ^. pileData))  (fromIntegral (subPile ^. pileCursor))
          (castPtr (pictData ^. pileData)) (fromIntegral (pictData ^. pileCursor))
          (castPtr (pictUsage ^. pileData)) (fromIntegral (pictUsage ^. pileCursor))
          bg (fromIntegral w) (fromIntegral h) frameCount
    forM_ (zip jobs [0..]) $ \(job, jobIndex) ->
      checkStatus ctx =<< c_rasterJob ctx
          (castPtr (job ^. rJShapePile . pileData)) (fromIntegral (job ^. rJShapePile . pileCursor))
          (castPtr (job ^. rJTilePile  . pileData)) (fromIntegral (job ^. rJTilePile  . pileCursor))
          (fromIntegral (job ^. rJColumnAllocation)) jobIndex
    case targetBuffer (params ^. rpTarget) of
      HostBitmapTarget outputPtr -> checkStatus ctx =<< c_frameEnd ctx outputPtr nullPtr
      GLTextureTarget _          -> error "GLTextureTarget not implemented"   -- as in the reference (:202-205)

-- | gudni_shape_entry, 32 bytes: what addShapeToTree receives (Raster/TileTree.hs:113) laid out for the C ABI —
-- the Shape GeoReference record of Raster/Types.hs:159-168 (tag, geoStart, strand count) followed by the bounding
-- box as four floats (left, top, right, bottom).  onShape (Raster/Serialize.hs:148-177) appends one to a Pile instead of
-- calling addShapeToTree.
data ShapeEntryRecord = ShapeEntryRecord
  { serTag     :: !ShapeTag
  , serGeoRef  :: !GeoReference
  , serBox     :: !(Box SubSpace)
  }

instance Storable ShapeEntryRecord where
  sizeOf    _ = 32
  alignment _ = 8
  peek ptr = do tag   <- peekByteOff ptr 0
                start <- peekByteOff ptr 8
                n     <- peekByteOff ptr 12 :: IO CUInt
                [l, t, r, b] <- mapM (\i -> peekByteOff ptr (16 + 4 * i)) [0 .. 3] :: IO [CFloat]
                return $ ShapeEntryRecord tag (GeoRef start (fromIntegral n))
                                          (makeBox (realToFrac l) (realToFrac t) (realToFrac r) (realToFrac b))
  poke ptr (ShapeEntryRecord tag (GeoRef start n) box) =
             do pokeByteOff ptr 0  tag
                pokeByteOff ptr 8  start
                pokeByteOff ptr 12 (fromIntegral n :: CUInt)
                pokeByteOff ptr 16 (realToFrac (box ^. leftSide)   :: CFloat)
                pokeByteOff ptr 20 (realToFrac (box ^. topSide)    :: CFloat)
                pokeByteOff ptr 24 (realToFrac (box ^. rightSide)  :: CFloat)
                pokeByteOff ptr 28 (realToFrac (box ^. bottomSide) :: CFloat)

-- | frame_begin with the frame constants of `params` (shared by the three entry points below).
withFrameInputs :: RasterParams token
                -> (Ptr CChar -> CSize -> Ptr CFloat -> CInt -> Ptr Word8 -> CSize -> Ptr () -> CInt -> Ptr CFloat -> CInt -> CInt -> IO a)
                -> IO a
withFrameInputs params k = do
    let geoPile  = params ^. rpGeometryState  . geoGeometryPile
        subPile  = params ^. rpSubstanceState . suSubstancePile
        Color' r g b a = colorComponents (params ^. rpSubstanceState . suBackgroundColor)
        V2 w h   = targetArea (params ^. rpTarget)
    (pictData, pictUsage) <- makePictData (params ^. rpSubstanceState . suPictureMapping)
                                          (params ^. rpSubstanceState . suPictureUsages)
    withArray [r, g, b, a] $ \bg ->
      k (castPtr (geoPile ^. pileData))   (fromIntegral (geoPile ^. pileCursor))
        (castPtr (subPile ^. pileData))   (fromIntegral (subPile ^. pileCursor))
        (castPtr (pictData ^. pileData))  (fromIntegral (pictData ^. pileCursor))
        (castPtr (pictUsage ^. pileData)) (fromIntegral (pictUsage ^. pileCursor))
        bg (fromIntegral w) (fromIntegral h)

-- | Level 2: skip the Haskell tile tree; hand the un-binned shape entries (tag, geoStart, strand
-- count, bounding box — what addShapeToTree receives, Raster/TileTree.hs:113) to the GPU binning.
-- Replaces buildTileTree / addShapeToTree / buildRasterJobs / queueRasterJobs (Application.hs:225-242) by one call.
queueRasterScene :: (MonadIO m, Show token)
                 => CInt -> RasterParams token -> Pile ShapeEntryRecord -> GeometryMonad m ()
queueRasterScene frameCount params entries = liftIO $ do
    let ctx = rasterCtx (params ^. rpDevice)
    withFrameInputs params $ \geo geoN sub subN pic picN use useN bg w h ->
      checkStatus ctx =<< c_frameBegin ctx geo geoN sub subN pic picN use useN bg w h frameCount
    checkStatus ctx =<< c_rasterScene ctx (castPtr (entries ^. pileData)) (fromIntegral (entries ^. pileCursor))
    case targetBuffer (params ^. rpTarget) of
      HostBitmapTarget outputPtr -> checkStatus ctx =<< c_frameEnd ctx outputPtr nullPtr
      GLTextureTarget _          -> error "GLTextureTarget not implemented"   -- as in the reference (:202-205)

-- | The same frame on every GPU of the box (a 16K x 16K canvas; a frame that fits one GPU is better left on one):
-- strips of whole root-tile rows, one per device, every device copying its rows into the HostBitmapTarget.
queueRasterSceneMulti :: (MonadIO m, Show token)
                      => Ptr GudniMulti -> CInt -> RasterParams token -> Pile ShapeEntryRecord -> GeometryMonad m ()
queueRasterSceneMulti multi frameCount params entries = liftIO $
    case targetBuffer (params ^. rpTarget) of
      GLTextureTarget _          -> error "GLTextureTarget not implemented"
      HostBitmapTarget outputPtr ->
        withFrameInputs params $ \geo geoN sub subN pic picN use useN bg w h -> do
          status <- c_multiFrame multi geo geoN sub subN pic picN use useN bg w h frameCount
                                 (castPtr (entries ^. pileData)) (fromIntegral (entries ^. pileCursor))
                                 nullPtr outputPtr nullPtr
          when (status /= 0) $ do msg <- peekCString =<< c_multiLastError multi
                                  error ("gudni_b200_multi_frame: " ++ show status ++ " " ++ msg)
