-- | Drop-in for Graphics.Gudni.OpenCL.Setup / Rasterizer (src/Graphics/Gudni/OpenCL/Setup.hs:102,
-- Rasterizer.hs:52-60).  SOURCE ONLY (no GHC in the build image).
module Graphics.Gudni.CUDA.Setup
  ( Rasterizer(..)
  , setupCUDA
  , checkStatus
  ) where

import Graphics.Gudni.CUDA.FFI
import Graphics.Gudni.OpenCL.Rasterizer (RasterSpec(..))

import Foreign.C.String (peekCString)
import Foreign.C.Types
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray)
import Foreign.Ptr
import Foreign.Storable

-- | The rasterizer handle: the context pointer replaces OpenCLState and the three CLKernels.
data Rasterizer = Rasterizer
  { rasterCtx  :: Ptr GudniCtx
  , rasterSpec :: RasterSpec
  }

-- | setupOpenCL's replacement: device pick + RasterSpec.  A null `want` asks for the canonical spec.
setupCUDA :: IO Rasterizer
setupCUDA =
  alloca $ \pCtx -> allocaArray 6 $ \got -> do
    status <- c_init (-1) nullPtr got pCtx
    if status /= 0 then error ("gudni_b200_init failed: " ++ show status) else do
      ctx <- peek pCtx
      [tile, threads, tilesPerCall, thresholds, strands, shapes] <- map fromIntegral <$> peekArray 6 got
      return $ Rasterizer ctx RasterSpec
        { _specMaxTileSize       = tile
        , _specThreadsPerTile    = threads
        , _specMaxTilesPerCall   = tilesPerCall
        , _specMaxThresholds     = thresholds
        , _specMaxStrandsPerTile = strands
        , _specMaxShapes         = shapes
        }

-- | The reference dies with `error` on failure (Setup.hs:116); keep that behaviour at the call sites.
checkStatus :: Ptr GudniCtx -> CInt -> IO ()
checkStatus _   0 = return ()
checkStatus ctx n = do msg <- peekCString =<< c_lastError ctx
                       error ("gudni_b200: " ++ show n ++ " " ++ msg)
