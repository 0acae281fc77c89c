{-# LANGUAGE ScopedTypeVariables #-}
-- | Level 3 of the C ABI (gudni_b200_raster_outlines): hand the scene to the library BEFORE serialisation.
-- SOURCE ONLY (no GHC in the build image).  Replaces, on the Haskell side, the body of onShape
-- (src/Graphics/Gudni/Raster/Serialize.hs:148-177): instead of transforming the outlines, boxing them,
-- culling against the canvas, building an Enclosure and appending it to the geometry pile — all of which the
-- library now does on the GPU — each shape only appends three small records.
--
-- Record layouts (include/gudni_b200.h):
--   gudni_outline_shape  32 B  tag :: Word64, firstOutline, nOutlines, firstTransform, nTransforms :: Word32, 8 B pad
--   gudni_outline         8 B  firstPair, nPairs :: Word32
--   gudni_curve_pair     16 B  onCurve.x, onCurve.y, offCurve.x, offCurve.y :: Float   (CurvePair, Figure/Outline.hs:55-57)
--   gudni_transform      16 B  kind :: Word32 (0 translate, 1 scale, 2 rotate), a, b :: Float, 4 B pad
module Graphics.Gudni.CUDA.Outlines
  ( OutlineState(..)
  , flattenTransformer
  , onShapeRaw
  , queueRasterOutlines
  ) where

import Graphics.Gudni.CUDA.FFI
import Graphics.Gudni.CUDA.Setup (Rasterizer(..), checkStatus)
import Graphics.Gudni.Figure
import Graphics.Gudni.Raster.Constants
import Graphics.Gudni.Raster.ShapeInfo
import Graphics.Gudni.Util.Pile

import Control.Monad.State
import Data.Word
import Foreign.C.Types
import Foreign.Ptr

-- | The four piles that cross the boundary, reused from frame to frame like geoGeometryPile.
data OutlineState = OutlineState
  { osShapes     :: Pile Word32   -- 8 words per shape
  , osOutlines   :: Pile Word32   -- 2 words per outline
  , osPairs      :: Pile CFloat   -- 4 floats per curve pair
  , osTransforms :: Pile Word32   -- 4 words per simple transformation (floats stored by bit pattern)
  }

-- | A Transformer as the list the library expects: outermost first, applied last to first.
-- applyTransformer (CombineTransform a b) = applyTransformer b . applyTransformer a
-- (Figure/Transformer.hs:100-105): a runs first, so it is the inner one and goes LAST.
flattenTransformer :: Transformer SubSpace -> [(Word32, Float, Float)]
flattenTransformer t = case t of
  Translate (Point2 x y) -> [(0, realToFrac x, realToFrac y)]
  Scale s                -> [(1, realToFrac s, 0)]
  Rotate a               -> let r = realToFrac (a ^. rad) :: Float   -- rotate (Figure/Angle.hs:52-53) multiplies
                            in  [(2, cos r, sin r)]                    -- by cos and sin of the angle
  CombineTransform a b   -> flattenTransformer b ++ flattenTransformer a

-- | What traverseShapeTree calls for every leaf (Raster/TraverseShapeTree.hs:73-80) instead of onShape.
-- Outlines that are placed many times (a glyph, a circle) can be appended once and referenced by index;
-- this version appends them per shape, which is what onShape's own serialisation costs today.
onShapeRaw :: SubstanceId -> SubstanceType -> Compound -> Transformer SubSpace -> [Outline SubSpace]
           -> StateT OutlineState IO ()
onShapeRaw substanceId substanceType combineType transformer outlines =
  do  st <- get
      let tag :: Word64
          tag = unShapeTag (makeShapeTag (ShapeInfo substanceType combineType substanceId))   -- Raster/ShapeInfo.hs:89-99
          firstOutline   = fromIntegral (osOutlines st ^. pileCursor) `div` 2
          firstTransform = fromIntegral (osTransforms st ^. pileCursor) `div` 4
          simple         = flattenTransformer transformer
      -- outlines and their pairs
      (outlinePile, pairPile) <- liftIO $ foldM appendOutline (osOutlines st, osPairs st) outlines
      -- transformer stack
      transformPile <- liftIO $ foldM appendTransform (osTransforms st) simple
      -- the shape record: tag (low word, high word), slices, padding
      shapePile <- liftIO $ foldM addToPile' (osShapes st)
                     [ fromIntegral tag, fromIntegral (tag `shiftR` 32)
                     , firstOutline, fromIntegral (length outlines)
                     , firstTransform, fromIntegral (length simple), 0, 0 ]
      put st { osShapes = shapePile, osOutlines = outlinePile, osPairs = pairPile, osTransforms = transformPile }
  where
    addToPile' pile x = fst <$> addToPile pile x
    appendOutline (outlinePile, pairPile) (Outline pairs) =
      do let firstPair = fromIntegral (pairPile ^. pileCursor) `div` 4 :: Word32
         pairPile' <- foldM (\p (CurvePair (Point2 ox oy) (Point2 cx cy)) ->
                               foldM addToPile' p (map realToFrac [ox, oy, cx, cy])) pairPile pairs
         outlinePile' <- foldM addToPile' outlinePile [firstPair, fromIntegral (length pairs)]
         return (outlinePile', pairPile')
    appendTransform pile (kind, a, b) = foldM addToPile' pile [kind, floatBits a, floatBits b, 0]
    floatBits = castFloatToWord32

-- | queueRasterJobs' counterpart: frame constants without a geometry pile, then the four piles.
-- Substances, pictures and the background go through c_frameBegin exactly as in
-- Graphics.Gudni.CUDA.CallKernels.queueRasterJobs.
queueRasterOutlines :: Rasterizer -> IO () {- ^ the c_frameBegin call of queueRasterJobs with nullPtr 0 for the geometry -}
                    -> OutlineState -> Ptr CUInt -> IO ()
queueRasterOutlines rasterizer frameBegin st outputPtr =
  do  let ctx = rasterCtx rasterizer
          count pile per = fromIntegral (pile ^. pileCursor) `div` per
      frameBegin
      checkStatus ctx =<< c_rasterOutlines ctx
          (castPtr (osShapes st ^. pileData))     (count (osShapes st) 8)
          (castPtr (osOutlines st ^. pileData))   (count (osOutlines st) 2)
          (castPtr (osPairs st ^. pileData))      (count (osPairs st) 4)
          (castPtr (osTransforms st ^. pileData)) (count (osTransforms st) 4)
      checkStatus ctx =<< c_frameEnd ctx outputPtr nullPtr
