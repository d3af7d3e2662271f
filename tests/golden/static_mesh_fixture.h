/* static-mesh fixture in the layout of a '3DS export' header (written for the tests of radiosity_b200; GL_T2F_N3F_V3F) */
class Cfixture {
	static const float p_object_0_vertices[8 * 8];
	static const unsigned int p_object_0_indices[3 * 12];
	static const float p_object_1_vertices[8 * 4];
	static const unsigned int p_object_1_indices[3 * 2];
};

const float Cfixture::p_object_0_vertices[8 * 8] = {
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.000000f, 0.000000f, 0.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 200.000000f, 0.000000f, 0.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 200.000000f, 200.000000f, 0.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.000000f, 200.000000f, 0.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.000000f, 0.000000f, 200.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 200.000000f, 0.000000f, 200.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 200.000000f, 200.000000f, 200.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.000000f, 200.000000f, 200.000000f
};
const unsigned int Cfixture::p_object_0_indices[3 * 12] = {
	0, 1, 2, 0, 2, 3, 4, 7, 6, 4, 6, 5, 0, 4, 5, 0, 5, 1, 1, 5, 6, 1, 6, 2, 2, 6, 7, 2, 7, 3, 3, 7, 4, 3, 4, 0
};
const Cfixture::TObject::TMatRange Cfixture::p_object_0_materials[1] = {
	{0, 0, 36}
};
const float Cfixture::p_object_1_vertices[8 * 4] = {
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 80.000000f, 199.000000f, 80.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 120.000000f, 199.000000f, 80.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 120.000000f, 199.000000f, 120.000000f,
	0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 80.000000f, 199.000000f, 120.000000f
};
const unsigned int Cfixture::p_object_1_indices[3 * 2] = {
	0, 1, 2, 0, 2, 3
};
const Cfixture::TObject::TMatRange Cfixture::p_object_1_materials[1] = {
	{1, 0, 6}
};
