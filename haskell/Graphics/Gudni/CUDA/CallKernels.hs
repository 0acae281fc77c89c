-- | Drop-in for Graphics.Gudni.OpenCL.CallKernels (src/Graphics/Gudni/OpenCL/CallKernels.hs):
-- same exported names and types, bodies call the C ABI instead of CLUtil.  SOURCE ONLY (no GHC in
-- the build image).  buildRasterJobs is unchanged from the reference (:244-255) and not repeated.
module Graphics.Gudni.CUDA.CallKernels
  ( queueRasterJobs
  , queueRasterScene
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
import Foreign.Ptr
import Foreign.C.Types

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

-- | Level 2: skip the Haskell tile tree; hand the un-binned shape entries (tag, geoStart, strand
-- count, bounding box — what addShapeToTree receives, Raster/TileTree.hs:113) to the GPU binning.
queueRasterScene :: (MonadIO m, Show token)
                 => CInt -> RasterParams token -> Pile ShapeEntryRecord -> GeometryMonad m ()
queueRasterScene frameCount params entries = error "see queueRasterJobs; replace the job loop by c_rasterScene"
