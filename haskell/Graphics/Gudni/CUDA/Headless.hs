-- | A headless output target: the frame ends in a host bitmap that is written to a file instead of being blitted to an
-- SDL window (the reference's only `handleOutput`s present the DrawTarget through SDL, Application.hs:84-101 /
-- Interface/InterfaceSDL.hs; its benchmarks and tests therefore need a display).  SOURCE ONLY (no GHC in the build
-- image).
--
--   * `withHostBitmap w h` allocates the pixels the rasterizer writes (what `prepareTarget` gets from SDL.lockTexture,
--     Interface/InterfaceSDL.hs) and page-locks them once (gudni_b200_host_register), so the read-back of every frame
--     is a direct DMA;
--   * `writePPM` stores the BGRA words (B | G<<8 | R<<16 | 0xFF<<24, Kernels.cl:842-844) as a binary P6 file;
--   * `handleOutputPPM` has the shape of `Model.handleOutput` for applications without a window.
module Graphics.Gudni.CUDA.Headless
  ( HostBitmap(..)
  , withHostBitmap
  , writePPM
  , handleOutputPPM
  ) where

import Graphics.Gudni.CUDA.FFI
import Graphics.Gudni.CUDA.Setup (Rasterizer(..), checkStatus)

import Control.Exception (bracket)
import Control.Monad (forM_)
import Data.Bits (shiftR, (.&.))
import qualified Data.ByteString as B
import qualified Data.ByteString.Builder as BB
import qualified Data.ByteString.Lazy as BL
import Foreign.C.Types
import Foreign.Marshal.Alloc (mallocBytes, free)
import Foreign.Marshal.Array (peekArray)
import Foreign.Ptr
import System.IO (withBinaryFile, IOMode(WriteMode))

-- | The pixels of one frame on the host: what `HostBitmapTarget` points at (Interface/DrawTarget.hs:33-35).
data HostBitmap = HostBitmap
  { hbWidth  :: !Int
  , hbHeight :: !Int
  , hbPixels :: !(Ptr CUInt)
  }

-- | Allocate, page-lock, run, unlock, free.
withHostBitmap :: Rasterizer -> Int -> Int -> (HostBitmap -> IO a) -> IO a
withHostBitmap rasterizer w h body =
    bracket acquire release (body . HostBitmap w h)
  where
    bytes   = w * h * 4
    ctx     = rasterCtx rasterizer
    acquire = do p <- mallocBytes bytes
                 checkStatus ctx =<< c_hostRegister ctx (castPtr p) (fromIntegral bytes)
                 checkStatus ctx =<< c_frameTargetHost ctx p      -- the kernels store into it; frame_end copies nothing
                 return p
    release p = do _ <- c_frameTargetHost ctx nullPtr
                   _ <- c_hostUnregister ctx (castPtr p)
                   free p

-- | Binary PPM (P6), top row first.
writePPM :: FilePath -> HostBitmap -> IO ()
writePPM path (HostBitmap w h pixels) =
    withBinaryFile path WriteMode $ \handle -> do
      B.hPut handle (BL.toStrict (BB.toLazyByteString (BB.string7 ("P6\n" ++ show w ++ " " ++ show h ++ "\n255\n"))))
      forM_ [0 .. h - 1] $ \row -> do
        ws <- peekArray w (pixels `plusPtr` (row * w * 4)) :: IO [CUInt]
        let rgb word = [ fromIntegral ((word `shiftR` 16) .&. 255)      -- R
                       , fromIntegral ((word `shiftR` 8) .&. 255)       -- G
                       , fromIntegral (word .&. 255) ]                  -- B
        B.hPut handle (B.pack (concatMap rgb ws))

-- | `handleOutput` for a model without a window: number the frames and write each to `prefix-NNNN.ppm`.
-- (In `Model s` the method also threads the InterfaceState; a headless application has none.)
handleOutputPPM :: String -> Int -> HostBitmap -> IO ()
handleOutputPPM prefix frame bitmap = writePPM (prefix ++ "-" ++ pad (show frame) ++ ".ppm") bitmap
  where pad s = replicate (4 - length s) '0' ++ s
